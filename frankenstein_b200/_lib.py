"""ctypes binding of the C-ABI library ``libfk_b200.so`` (see include/fk_b200.h).

The prototypes are read from the header itself, so the Python side cannot drift from the ABI.
There is no fallback: if the library is missing or was not built for the device in use, every
entry point raises.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# FK_LIB_PATH: load an experiment build of the same library (frankenstein_b200.build --variant=...) for kernel diagnosis
LIB_PATH = os.environ.get("FK_LIB_PATH") or os.path.join(_HERE, "libfk_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "fk_b200.h")

_CTYPE = {
    "int": ctypes.c_int, "float": ctypes.c_float, "long long": ctypes.c_longlong,
    "unsigned int": ctypes.c_uint, "void": None,
}


class FkError(RuntimeError):
    pass


def parse_header(path: str = HEADER_PATH) -> Dict[str, Tuple[str, List[str]]]:
    """{name: (return type, [argument types])} for every prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = "\n".join(l for l in text.splitlines() if not l.lstrip().startswith("#") and 'extern "C"' not in l)
    protos = {}
    for m in re.finditer(r"([\w\s\*]+?)\b(fk_\w+)\s*\(([^)]*)\)\s*;", text):
        ret = " ".join(m.group(1).split())
        args = []
        for a in m.group(3).split(","):
            a = " ".join(a.split())
            if a in ("void", ""):
                continue
            if "*" in a:
                args.append("ptr")
            else:
                args.append(" ".join(a.split()[:-1]).replace("const ", ""))
        protos[m.group(2)] = (ret, args)
    return protos


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FkError(f"{LIB_PATH} is missing: run `python -m frankenstein_b200.build` (there is no CPU/PyTorch "
                      f"fallback for the sm_100a kernels)")
    L = ctypes.CDLL(LIB_PATH)
    for name, (ret, args) in parse_header().items():
        fn = getattr(L, name)  # AttributeError if the header declares a symbol the library lacks
        fn.argtypes = [ctypes.c_void_p if a == "ptr" else _CTYPE[a] for a in args]
        fn.restype = ctypes.c_char_p if "char" in ret else (_CTYPE[ret.replace("const ", "")] if "*" not in ret else ctypes.c_void_p)
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().fk_last_error()
        raise FkError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise FkError("frankenstein_b200 kernels run on a B200 (sm_100a) only; got a CPU tensor "
                          "(no CPU fallback exists)")


_device_checked = False


def require_device() -> None:
    global _device_checked
    if _device_checked:
        return
    if not torch.cuda.is_available():
        raise FkError("no CUDA device: frankenstein_b200 has no CPU fallback")
    if not lib().fk_device_ok():
        raise FkError("libfk_b200.so is compiled for sm_100a only and the current device is not compute capability 10.x")
    _device_checked = True


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(lib().fk_launch_count())


def reset_launch_count() -> None:
    lib().fk_reset_launch_count()


DTYPE_CODE = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


# ------------------------------------------------------------------------------------------------
# optional per-kernel timing (bench.py): CUDA events recorded on the launching stream around a call
# ------------------------------------------------------------------------------------------------
class KernelTimer:
    """`with timed("name", work): ...` records two events on the current stream when enabled; `summary()`
    synchronises and returns {name: (launches, total_ms, total_work)}.  Disabled = zero overhead."""

    def __init__(self):
        self.enabled = False
        self.records = []

    def reset(self):
        self.records = []

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, a, b, work in self.records:
            n, ms, w = out.get(name, (0, 0.0, 0.0))
            out[name] = (n + 1, ms + a.elapsed_time(b), w + work)
        return out


TIMER = KernelTimer()


class timed:
    def __init__(self, name, work=0.0):
        self.name, self.work = name, work

    def __enter__(self):
        if TIMER.enabled:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if TIMER.enabled:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            TIMER.records.append((self.name, self.a, b, self.work))
        return False

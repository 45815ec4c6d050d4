"""ctypes binding of the C-ABI library ``libfk_b200.so`` (see include/fk_b200.h).

The prototypes are read from the header itself, so the Python side cannot drift from the ABI.
There is no fallback: if the library is missing or was not built for the device in use, every
entry point raises.
"""
from __future__ import annotations

import ctypes
import functools
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
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise FkError("frankenstein_b200 kernels run on a B200 (sm_100a) only; got a CPU tensor "
                          "(no CPU fallback exists)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise FkError(f"tensors of one kernel call live on different devices ({dev} and {t.device})")


def on_tensor_device(fn):
    """Run `fn` with the device of its first CUDA tensor argument current: the C entry points launch on the current
    device and on the stream `stream()` reports for it, so a model that lives on cuda:1 works without the caller having
    to `torch.cuda.set_device(1)` first (host-side caches of the library are keyed by the device ordinal)."""
    @functools.wraps(fn)
    def wrapper(*args, **kw):
        idx = None
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                idx = a.device.index
                break
        if idx is None or idx == torch.cuda.current_device():
            return fn(*args, **kw)
        with torch.cuda.device(idx):
            return fn(*args, **kw)
    return wrapper


_device_checked = set()


def require_device() -> None:
    """The CURRENT device must be a compute-capability 10.x part (checked once per device ordinal)."""
    if not torch.cuda.is_available():
        raise FkError("no CUDA device: frankenstein_b200 has no CPU fallback")
    idx = torch.cuda.current_device()
    if idx in _device_checked:
        return
    if not lib().fk_device_ok():
        raise FkError("libfk_b200.so is compiled for sm_100a only and the current device is not compute capability 10.x")
    _device_checked.add(idx)


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def stream() -> int:
    """cudaStream_t of torch's current stream on the current device (ABI calls run under `on_tensor_device`)."""
    return torch.cuda.current_stream().cuda_stream


# caller-owned counter words of the ABI (work hand-out of the persistent kernels, last-block reductions): zero on entry,
# put back to zero by the kernel.  One set per (device, stream): launches on one stream are ordered, so they can share
# the words; launches on different streams get different ones and may overlap freely.
COUNTER_WORDS = 32
CTR_VQ_FINISH, CTR_MASKED_L1, CTR_ATTN_FWD, CTR_ATTN_BWD, CTR_GEMM = 0, 1, 2, 4, 8
_counter_sets = {}


def counters(slot: int = 0) -> int:
    """device pointer to counter word `slot` of the current (device, stream)."""
    key = (torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
    t = _counter_sets.get(key)
    if t is None:
        t = torch.zeros(COUNTER_WORDS, device=torch.device("cuda", key[0]), dtype=torch.int32)
        _counter_sets[key] = t
    return t.data_ptr() + 4 * slot


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
        self.only = None          # tuple of name prefixes to record (None = every timed call)
        self.records = []

    def reset(self):
        self.records = []

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, a, b, work in self.records:
            n, ms, w = out.get(name, (0, 0.0, 0.0))
            out[name] = (n + 1, ms + a.elapsed_time(b), w + float(work))      # (work may be a 0-dim device tensor)
        return out


TIMER = KernelTimer()


class timed:
    def __init__(self, name, work=0.0):
        self.name, self.work = name, work

    def __enter__(self):
        self.on = TIMER.enabled and (TIMER.only is None or self.name.startswith(TIMER.only))
        if self.on:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if self.on:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            TIMER.records.append((self.name, self.a, b, self.work))
        return False

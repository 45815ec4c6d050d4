"""Input side of the hot path on the device -- host-side mirror of the reference's ``utils/data_utils.py`` signal
preprocessing, same function names and argument meaning:

* ``process_signal(voltage_list, spikes_list, block_list)``        utils/data_utils.py:115-156
* ``z_score_per_block_scaling(brain_list, idx_list)``               utils/data_utils.py:78-109
* ``pad_truncate_brain_list(brain_list, max_length)``               utils/data_utils.py:243-267
* ``make_batch(...)``: the three steps above plus the float32 cast of ``BrainDataset.__getitem__`` (:335-344) fused into
  ONE normalisation pass that writes the padded ``[n_trials, max_length, C]`` batch directly (fp32, or bf16 for the
  encoder's patch GEMM) -- the z-scored, smoothed and padded intermediates never exist;
* ``DevicePrefetcher``: pinned-memory host batches copied one step ahead on a copy stream (the reference's 3-worker NumPy
  DataLoader, utils/train_utils.py:77-83, is the bottleneck above ~10 k trials/s).

Trials are ragged ([T_i, C] each).  They are packed back to back on the device (one H2D copy per feature kind) and handed
to ``csrc/input_ops.cu`` with their row offsets and dense block ids.  There is no CPU path: the kernels run on a B200, the
lists may hold numpy arrays or tensors on any device.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from ._lib import FkError, check, lib, ptr, require_device, stream

ZERO_EXACT, ZERO_SKLEARN = 0, 1


def _device(device) -> torch.device:
    dev = torch.device(device if device is not None else "cuda")
    if dev.type != "cuda":
        raise FkError("frankenstein_b200.data_pipeline runs on a B200 only (no CPU fallback)")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _pack(arrays: Sequence, dev: torch.device) -> Tuple[torch.Tensor, List[int]]:
    """[T_i, C] arrays -> one fp32 device tensor [sum_T, C] (a single pinned staging buffer and copy when the trials come
    from the host) and the list of lengths."""
    lengths = [int(a.shape[0]) for a in arrays]
    if len(arrays) == 0:
        raise FkError("empty trial list")
    C = int(arrays[0].shape[1])
    if any(a.ndim != 2 or a.shape[1] != C for a in arrays):
        raise FkError("every trial must be [T_i, C] with the same channel count")
    if all(isinstance(a, torch.Tensor) and a.is_cuda for a in arrays):
        return torch.cat([a.to(dev, torch.float32) for a in arrays], dim=0).contiguous(), lengths
    total = sum(lengths)
    host = torch.empty(total, C, dtype=torch.float32, pin_memory=True)
    o = 0
    for a, n in zip(arrays, lengths):
        host[o:o + n] = torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a.cpu(), dtype=torch.float32)
        o += n
    return host.to(dev, non_blocking=True), lengths


def _dense_blocks(block_list) -> Tuple[np.ndarray, int]:
    ids = np.asarray(block_list.cpu() if isinstance(block_list, torch.Tensor) else block_list)
    _, dense = np.unique(ids, return_inverse=True)
    return dense.astype(np.int32).reshape(-1), int(dense.max()) + 1


class PackedTrials:
    """Ragged trials on the device: ``volt [sum_T, C1]``, ``spk [sum_T, C2] | None``, ``offsets [n + 1]`` int64,
    ``block_id [n]`` int32 (dense), plus the per-block statistics once computed."""

    def __init__(self, voltage_list, spikes_list, block_list, device=None):
        dev = _device(device)
        with torch.cuda.device(dev):
            require_device()
            self.volt, self.lengths = _pack(voltage_list, dev)
            self.spk = None
            if spikes_list is not None:
                self.spk, l2 = _pack(spikes_list, dev)
                if l2 != self.lengths:
                    raise FkError("voltage and spike trials differ in length")
            if len(block_list) != len(self.lengths):
                raise FkError("block_list must name one block per trial")
            self.C1 = self.volt.shape[1]
            self.C2 = 0 if self.spk is None else self.spk.shape[1]
            if self.C1 % 4 or self.C2 % 4:
                raise FkError("channel counts must be multiples of 4 (the kernels move float4 channel quads)")
            dense, self.n_blocks = _dense_blocks(block_list)
            off = np.zeros(len(self.lengths) + 1, dtype=np.int64)
            off[1:] = np.cumsum(self.lengths)
            self.offsets = torch.from_numpy(off).to(dev)
            self.block_id = torch.from_numpy(dense).to(dev)
            self.device = dev
            self.mean = self.std = None

    @property
    def n_trials(self) -> int:
        return len(self.lengths)

    @property
    def C(self) -> int:
        return self.C1 + self.C2

    def block_stats(self, zero_policy: int = ZERO_EXACT):
        """per block and channel: mean and population std over the concatenation of the block's trials (two passes:
        sum -> mean, squared deviations -> std; fp64 accumulation in a fixed order), std == 0 -> 1."""
        with torch.cuda.device(self.device):
            n, C, nb = self.n_trials, self.C, self.n_blocks
            part = torch.empty(n, C, device=self.device, dtype=torch.float64)
            mean_d = torch.empty(nb, C, device=self.device, dtype=torch.float64)
            self.mean = torch.empty(nb, C, device=self.device, dtype=torch.float32)
            self.std = torch.empty(nb, C, device=self.device, dtype=torch.float32)
            L, s = lib(), stream()
            check(L.fk_input_trial_moments(ptr(self.volt), ptr(self.spk), ptr(self.offsets), ptr(self.block_id), n, self.C1,
                                           self.C2, None, ptr(part), s), "fk_input_trial_moments")
            check(L.fk_input_block_reduce(ptr(part), ptr(self.offsets), ptr(self.block_id), n, nb, C, 0, zero_policy,
                                          ptr(mean_d), ptr(self.mean), None, s), "fk_input_block_reduce")
            check(L.fk_input_trial_moments(ptr(self.volt), ptr(self.spk), ptr(self.offsets), ptr(self.block_id), n, self.C1,
                                           self.C2, ptr(mean_d), ptr(part), s), "fk_input_trial_moments")
            check(L.fk_input_block_reduce(ptr(part), ptr(self.offsets), ptr(self.block_id), n, nb, C, 1, zero_policy,
                                          None, None, ptr(self.std), s), "fk_input_block_reduce")
        return self.mean, self.std

    def normalize(self, T_out: int, smooth: bool, out_dtype=torch.float32, zero_policy: int = ZERO_EXACT) -> torch.Tensor:
        """[n_trials, T_out, C]: z-score (+ Gaussian smoothing over each trial's own bins), zero padded / truncated."""
        if out_dtype not in (torch.float32, torch.bfloat16):
            raise FkError("out_dtype must be float32 or bfloat16")
        if self.mean is None:
            self.block_stats(zero_policy)
        with torch.cuda.device(self.device):
            out = torch.empty(self.n_trials, T_out, self.C, device=self.device, dtype=out_dtype)
            check(lib().fk_input_normalize(ptr(self.volt), ptr(self.spk), ptr(self.offsets), ptr(self.block_id),
                                           ptr(self.mean), ptr(self.std), self.n_trials, self.C1, self.C2, T_out,
                                           1 if smooth else 0, ptr(out), 0 if out_dtype == torch.float32 else 1, stream()),
                  "fk_input_normalize")
        return out


def make_batch(voltage_list, spikes_list, block_list, max_length: int, smooth: bool = True, out_dtype=torch.float32,
               device=None) -> torch.Tensor:
    """process_signal -> pad_truncate_brain_list -> float32 -> stack, as one device pass: [n_trials, max_length, C1 + C2]."""
    return PackedTrials(voltage_list, spikes_list, block_list, device).normalize(max_length, smooth, out_dtype)


def process_signal(voltage_list, spikes_list, block_list, device=None) -> List[torch.Tensor]:
    """utils/data_utils.py:115-156: list of [T_i, C1 + C2] fp32 device tensors (views of one padded buffer)."""
    pk = PackedTrials(voltage_list, spikes_list, block_list, device)
    full = pk.normalize(max(pk.lengths), True)
    return [full[i, :n] for i, n in enumerate(pk.lengths)]


def z_score_per_block_scaling(brain_list, idx_list, device=None) -> List[torch.Tensor]:
    """utils/data_utils.py:78-109 (StandardScaler per block: population std, near-zero scale -> 1)."""
    pk = PackedTrials(brain_list, None, idx_list, device)
    full = pk.normalize(max(pk.lengths), False, zero_policy=ZERO_SKLEARN)
    return [full[i, :n] for i, n in enumerate(pk.lengths)]


def pad_truncate_brain_list(brain_list, max_length: int, device=None) -> List[torch.Tensor]:
    """utils/data_utils.py:243-267 on device tensors: zero-pad at the end or truncate to max_length bins."""
    dev = _device(device)
    out = []
    for x in brain_list:
        x = torch.as_tensor(x).to(dev)
        T = x.shape[0]
        if T >= max_length:
            out.append(x[:max_length])
        else:
            p = x.new_zeros(max_length, x.shape[1])
            p[:T] = x
            out.append(p)
    return out


class DevicePrefetcher:
    """Iterates over host batches (tensors or tuples of tensors), staging each in pinned memory and copying it to the
    device on a side stream one step ahead of the consumer; the consumer's stream waits on the copy's event, never the
    host.  Replaces the synchronous `.to(device)` of the reference's training loop (utils/train_utils.py:131-134)."""

    def __init__(self, batches, device=None, depth: int = 2):
        self.dev = _device(device)
        self.it = iter(batches)
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.depth = max(1, depth)
        self.queue = []
        self.h2d_bytes = 0
        for _ in range(self.depth):
            self._enqueue()

    def _enqueue(self):
        try:
            item = next(self.it)
        except StopIteration:
            return
        single = isinstance(item, torch.Tensor)
        host = [item] if single else list(item)
        with torch.cuda.stream(self.copy_stream):
            dev_t = []
            for t in host:
                if isinstance(t, torch.Tensor):
                    if not t.is_cuda:
                        if not t.is_pinned():
                            t = t.pin_memory()
                        self.h2d_bytes += t.numel() * t.element_size()
                    dev_t.append(t.to(self.dev, non_blocking=True))
                else:
                    dev_t.append(t)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.queue.append((dev_t[0] if single else tuple(dev_t), ev))

    def __iter__(self):
        return self

    def __next__(self):
        if not self.queue:
            raise StopIteration
        item, ev = self.queue.pop(0)
        torch.cuda.current_stream(self.dev).wait_event(ev)
        for t in ([item] if isinstance(item, torch.Tensor) else item):
            if isinstance(t, torch.Tensor):
                t.record_stream(torch.cuda.current_stream(self.dev))
        self._enqueue()
        return item

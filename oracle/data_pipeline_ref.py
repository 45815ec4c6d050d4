"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the input side of the hot path: numpy restatement of the reference's
signal preprocessing, `utils/data_utils.py`.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import
this; the product path (frankenstein_b200/data_pipeline.py) runs on the GPU and never falls back to it.

Pinned: tests/golden/data_pipeline.npz holds outputs of the UNMODIFIED reference functions (generated in the build
container by scripts/make_golden_pipeline.py, which imports /root/reference/utils/data_utils.py directly -- it needs
only numpy / scipy / scikit-learn); tests/test_oracle_cpu.py checks this restatement against them.

Restated functions (reference file:line):
  z_score_per_block_scaling   utils/data_utils.py:78-109   per block: StandardScaler().fit(concat of the block's trials)
                                                           -> (x - mean) / scale, scale = std (ddof 0), 1 where std == 0
  process_signal              utils/data_utils.py:115-156  [voltage | spikes] channel concat, per-block mean / std
                                                           (std == 0 -> 1), z-score, gaussian_filter1d(sigma=1, axis=0)
  pad_truncate_brain_list     utils/data_utils.py:243-267  zero-pad at the end / truncate to max_length bins
  BrainDataset.__getitem__    utils/data_utils.py:335-344  .astype(np.float32)
"""
from __future__ import annotations

import numpy as np

GAUSS_SIGMA = 1.0
GAUSS_RADIUS = 4          # scipy: int(truncate * sigma + 0.5) with truncate = 4.0


def gaussian_weights(sigma: float = GAUSS_SIGMA, radius: int = GAUSS_RADIUS) -> np.ndarray:
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, order=0, radius): exp(-x^2 / (2 sigma^2)) normalised to sum 1."""
    x = np.arange(-radius, radius + 1, dtype=np.float64)
    w = np.exp(-0.5 / (sigma * sigma) * x * x)
    return w / w.sum()


def reflect_index(i: int, n: int) -> int:
    """scipy mode='reflect' (half-sample symmetric: d c b a | a b c d | d c b a)."""
    if n == 1:
        return 0
    period = 2 * n
    i %= period
    if i < 0:
        i += period
    return i if i < n else period - 1 - i


def gaussian_filter_time(x: np.ndarray) -> np.ndarray:
    """scipy.ndimage.gaussian_filter1d(x, sigma=1, axis=0) (mode='reflect', truncate=4.0) on [T, C]; float64
    accumulation, result in x's dtype."""
    T = x.shape[0]
    w = gaussian_weights()
    out = np.zeros(x.shape, dtype=np.float64)
    xd = x.astype(np.float64)
    for k in range(-GAUSS_RADIUS, GAUSS_RADIUS + 1):
        idx = np.array([reflect_index(t + k, T) for t in range(T)])
        out += w[k + GAUSS_RADIUS] * xd[idx]
    return out.astype(x.dtype)


def block_stats(brain_list, block_list):
    """per block: column mean and std (ddof 0) over the concatenation of the block's trials; std == 0 -> 1."""
    stats = {}
    block_list = np.asarray(block_list)
    for blk in np.unique(block_list):
        cat = np.concatenate([brain_list[i] for i in np.nonzero(block_list == blk)[0]], axis=0)
        mean = cat.mean(axis=0)
        std = cat.std(axis=0)
        std = np.where(std == 0, 1.0, std)
        stats[blk] = (mean, std)
    return stats


def z_score_per_block_scaling(brain_list, idx_list):
    """utils/data_utils.py:78-109 (StandardScaler: population std; near-zero scale -> 1)."""
    st = block_stats(brain_list, idx_list)
    out = []
    for x, blk in zip(brain_list, idx_list):
        mean, std = st[blk]
        # sklearn's _handle_zeros_in_scale also maps scales below 10 * eps to 1
        eps = 10 * np.finfo(std.dtype).eps
        std = np.where(std < eps, 1.0, std)
        out.append((x - mean) / std)
    return out


def process_signal(voltage_list, spikes_list, block_list):
    """utils/data_utils.py:115-156."""
    n = len(block_list)
    cat = [np.concatenate([voltage_list[i], spikes_list[i]], axis=1) for i in range(n)]
    st = block_stats(cat, block_list)
    out = np.empty(n, dtype=object)
    for i in range(n):
        mean, std = st[np.asarray(block_list)[i]]
        out[i] = gaussian_filter_time((cat[i] - mean[None]) / std[None])
    return out


def pad_truncate_brain_list(brain_list, max_length):
    """utils/data_utils.py:243-267."""
    out = []
    for x in brain_list:
        T = x.shape[0]
        out.append(x[:max_length] if T > max_length else np.pad(x, ((0, max_length - T), (0, 0)), mode="constant"))
    return out


def make_batch(voltage_list, spikes_list, block_list, max_length, smooth=True, exact=False):
    """process_signal -> pad_truncate_brain_list -> astype(float32) -> stack: the [n, max_length, C] batch the trainer's
    DataLoader hands to the model (utils/data_utils.py:115-156, :243-267, :335-344).
    exact=True evaluates the same formulas in float64 (the reference works in the arrays' own float32, where numpy's
    axis-0 mean / std add rows one by one: at ~50 k bins per block its statistics carry ~1e-5 relative rounding error)."""
    if exact:
        voltage_list = [np.asarray(v, dtype=np.float64) for v in voltage_list]
        spikes_list = [np.asarray(v, dtype=np.float64) for v in spikes_list]
    if smooth:
        proc = process_signal(voltage_list, spikes_list, block_list)
    else:
        cat = [np.concatenate([v, s], axis=1) for v, s in zip(voltage_list, spikes_list)]
        st = block_stats(cat, block_list)
        proc = [(c - st[b][0][None]) / st[b][1][None] for c, b in zip(cat, block_list)]
    return np.stack([p.astype(np.float32) for p in pad_truncate_brain_list(list(proc), max_length)])

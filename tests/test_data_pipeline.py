"""Input pipeline (utils/data_utils.py): the numpy oracle against fixtures produced by the UNMODIFIED
reference functions (CPU), and the device kernels against the oracle and the same fixtures (GPU).  Floating point: the
reference works in float32 numpy with pairwise sums, the kernels accumulate in fp64 in a fixed order and divide in fp32 --
stated tolerance 2e-5 absolute on O(1) z-scores (bf16 output: 1e-2)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "data_pipeline.npz")
ATOL = 2e-5


def _golden():
    g = np.load(GOLD)
    n = len(g["lengths"])
    volt = [g[f"volt{i}"] for i in range(n)]
    spk = [g[f"spk{i}"] for i in range(n)]
    return g, volt, spk, g["blocks"], int(g["max_len"])


def test_oracle_matches_reference_fixture():
    from oracle import data_pipeline_ref as R
    g, volt, spk, blocks, max_len = _golden()
    proc = R.process_signal(volt, spk, blocks)
    for i, p in enumerate(proc):
        assert p.shape == g[f"proc{i}"].shape
        np.testing.assert_allclose(p, g[f"proc{i}"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(R.make_batch(volt, spk, blocks, max_len), g["batch"], rtol=0, atol=1e-6)
    cat = [np.concatenate([v, s], axis=1) for v, s in zip(volt, spk)]
    z = np.stack([p.astype(np.float32) for p in R.pad_truncate_brain_list(R.z_score_per_block_scaling(cat, list(blocks)), max_len)])
    np.testing.assert_allclose(z, g["zscore_batch"], rtol=0, atol=5e-6)
    # constant channels (std == 0 -> 1) come out as exact zeros, padding rows too
    assert np.all(g["batch"][0, 33:] == 0) and np.all(R.make_batch(volt, spk, blocks, max_len)[0, 33:] == 0)


def test_oracle_gaussian_matches_scipy():
    import scipy.ndimage
    from oracle import data_pipeline_ref as R
    rng = np.random.default_rng(0)
    for T in (1, 2, 3, 5, 9, 40):
        x = rng.standard_normal((T, 8)).astype(np.float32)
        np.testing.assert_allclose(R.gaussian_filter_time(x), scipy.ndimage.gaussian_filter1d(x, sigma=1, axis=0), rtol=0, atol=1e-6)


def test_host_side_argument_checks():
    from frankenstein_b200 import data_pipeline as dp
    from frankenstein_b200._lib import FkError
    with pytest.raises(FkError):
        dp.make_batch([np.zeros((4, 8), np.float32)], [np.zeros((4, 8), np.float32)], [0], 8, device="cpu")


@pytest.mark.gpu
def test_make_batch_matches_reference_fixture_and_oracle():
    from frankenstein_b200 import data_pipeline as dp
    g, volt, spk, blocks, max_len = _golden()
    out = dp.make_batch(volt, spk, blocks, max_len)
    assert out.shape == g["batch"].shape and out.dtype == torch.float32 and out.is_cuda
    np.testing.assert_allclose(out.cpu().numpy(), g["batch"], rtol=0, atol=ATOL)
    assert torch.all(out[0, 33:] == 0) and torch.all(out[4, 1:] == 0)          # padding is exact zeros
    proc = dp.process_signal(volt, spk, blocks)
    for i, p in enumerate(proc):
        np.testing.assert_allclose(p.cpu().numpy(), g[f"proc{i}"], rtol=0, atol=ATOL)
    cat = [np.concatenate([v, s], axis=1) for v, s in zip(volt, spk)]
    zs = dp.pad_truncate_brain_list(dp.z_score_per_block_scaling(cat, list(blocks)), max_len)
    np.testing.assert_allclose(torch.stack(zs).cpu().numpy(), g["zscore_batch"], rtol=0, atol=ATOL)
    bf = dp.make_batch(volt, spk, blocks, max_len, out_dtype=torch.bfloat16)
    np.testing.assert_allclose(bf.float().cpu().numpy(), g["batch"], rtol=1e-2, atol=1e-2)
    # run-to-run identical (fixed-order fp64 reductions)
    assert torch.equal(out, dp.make_batch(volt, spk, blocks, max_len))


@pytest.mark.gpu
@pytest.mark.parametrize("smooth", [True, False])
def test_make_batch_full_size_against_oracle(smooth):
    """cfg-4 batch shape: 128 ragged trials x 256 + 256 channels -> [128, 512, 512], 12 recording blocks."""
    from frankenstein_b200 import data_pipeline as dp
    from oracle import data_pipeline_ref as R
    rng = np.random.default_rng(7)
    n, ch = 128, 256
    lengths = rng.integers(300, 640, size=n)
    lengths[:3] = (512, 1, 700)
    blocks = rng.integers(0, 12, size=n) * 3
    volt = [(rng.standard_normal((T, ch)) * (1 + b) + 5).astype(np.float32) for T, b in zip(lengths, blocks)]
    spk = [rng.poisson(2.0, size=(T, ch)).astype(np.float32) for T in lengths]
    out = dp.make_batch(volt, spk, blocks, 512, smooth=smooth)
    # the float32 reference arithmetic (numpy adds ~50 k rows per block one by one in float32: its block statistics carry
    # ~1e-5 relative error) within 1e-3; the same formulas evaluated in float64 within 5e-6 -- the kernels accumulate in fp64
    ref = R.make_batch(volt, spk, blocks, 512, smooth=smooth)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=0, atol=1e-3)
    exact = R.make_batch(volt, spk, blocks, 512, smooth=smooth, exact=True)
    np.testing.assert_allclose(out.cpu().numpy(), exact, rtol=0, atol=5e-6)
    # size-independent properties: per-block statistics of the un-smoothed z-scores are (0, 1); padding is zero
    if not smooth:
        pk = dp.PackedTrials(volt, spk, blocks, "cuda")
        z = pk.normalize(int(lengths.max()), False)
        for b in np.unique(blocks)[:4]:
            rows = torch.cat([z[i, :lengths[i]] for i in np.nonzero(blocks == b)[0]]).double()
            assert rows.mean(0).abs().max() < 1e-5 and (rows.std(0, unbiased=False) - 1).abs().max() < 1e-4
    for i in range(n):
        if lengths[i] < 512:
            assert torch.all(out[i, lengths[i]:] == 0)


@pytest.mark.gpu
def test_device_prefetcher_delivers_batches_in_order():
    from frankenstein_b200.data_pipeline import DevicePrefetcher
    batches = [(torch.full((4, 8), float(i)), torch.arange(4) + i) for i in range(5)]
    got = list(DevicePrefetcher(batches, "cuda"))
    assert len(got) == 5
    for i, (x, y) in enumerate(got):
        assert x.is_cuda and torch.all(x == i) and torch.equal(y.cpu(), torch.arange(4) + i)

"""Generate tests/golden/data_pipeline.npz from the UNMODIFIED reference utils/data_utils.py (build container only:
/root/reference does not exist on the GPU box).  The module needs numpy / scipy / scikit-learn / torch only, so it is
imported as it is -- no shims."""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("FK_REFERENCE", "/root/reference")


def load_data_utils():
    spec = importlib.util.spec_from_file_location("ref_data_utils", os.path.join(REF, "utils", "data_utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def synth_trials(rng, n, c_half, lengths, blocks):
    """spike-power-like (positive, per-channel gain) and threshold-crossing-like (small integers) features; one channel
    of each kind is constant inside every block (std == 0 -> scale 1)."""
    volt, spk = [], []
    for i in range(n):
        T = lengths[i]
        gain = 1.0 + 0.1 * blocks[i] + rng.random(c_half)
        v = (rng.standard_normal((T, c_half)) * gain + 3.0 * gain).astype(np.float32)
        s = rng.poisson(1.5 + 0.2 * blocks[i], size=(T, c_half)).astype(np.float32)
        v[:, 3] = 7.0 + blocks[i]
        s[:, 5] = 0.0
        volt.append(v)
        spk.append(s)
    return volt, spk


def main():
    du = load_data_utils()
    rng = np.random.default_rng(1234)
    n, c_half, max_len = 7, 16, 40
    lengths = [33, 40, 57, 12, 1, 40, 45]                 # shorter, equal, longer than max_len, and a 1-bin trial
    blocks = np.array([2, 2, 5, 5, 9, 2, 5])              # unsorted block ids, one single-trial block
    volt, spk = synth_trials(rng, n, c_half, lengths, blocks)
    proc = du.process_signal(volt, spk, blocks)
    padded = du.pad_truncate_brain_list(list(proc), max_len)
    batch = np.stack([p.astype(np.float32) for p in padded])
    cat = [np.concatenate([v, s], axis=1) for v, s in zip(volt, spk)]
    zs = du.z_score_per_block_scaling(cat, list(blocks))
    zs_padded = np.stack([p.astype(np.float32) for p in du.pad_truncate_brain_list(zs, max_len)])
    out = {"lengths": np.array(lengths), "blocks": blocks, "max_len": np.array(max_len), "batch": batch,
           "zscore_batch": zs_padded}
    for i in range(n):
        out[f"volt{i}"], out[f"spk{i}"], out[f"proc{i}"] = volt[i], spk[i], proc[i].astype(np.float64)
    path = os.path.join(ROOT, "tests", "golden", "data_pipeline.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")




def gpt2_prefix_fixture():
    """models/gpt2_model.py:178-196 -- what the UNMODIFIED reference GPT feeds its first block for (idx, prefix), captured
    with a hook on transformer.drop, plus the gradients the whole decoder sends back through the hand-off."""
    import contextlib
    import io

    import torch
    spec = importlib.util.spec_from_file_location("ref_gpt2_model", os.path.join(REF, "models", "gpt2_model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(3)
    with contextlib.redirect_stdout(io.StringIO()):
        gpt = mod.GPT(mod.GPTConfig(block_size=48, vocab_size=64, n_layer=1, n_head=2, n_embd=32, dropout=0.0, bias=True))
    B, Tc, T = 3, 5, 7
    idx = torch.randint(0, 64, (B, T))
    idx[0, 1] = idx[0, 0]                                   # a repeated token: its wte gradient rows add up
    prefix = torch.randn(B, Tc, 32, requires_grad=True)
    cap = {}

    def hook(m, inp):
        cap["x0"] = inp[0]
        inp[0].retain_grad()

    h = gpt.transformer.drop.register_forward_pre_hook(hook)
    loss, _ = gpt(idx, prefix=prefix, targets=idx)
    loss.backward()
    h.remove()
    out = {"idx": idx, "prefix": prefix.detach(), "wte": gpt.transformer.wte.weight.detach().clone(),
           "wpe": gpt.transformer.wpe.weight.detach().clone(), "x0": cap["x0"].detach().clone(), "g": cap["x0"].grad.clone(),
           "dprefix": prefix.grad.clone(), "dwpe": gpt.transformer.wpe.weight.grad.clone()}
    path = os.path.join(ROOT, "tests", "golden", "gpt2_prefix.pt")
    torch.save(out, path)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
    gpt2_prefix_fixture()

"""CPU oracle (TEST INFRASTRUCTURE ONLY): the hand-off of the perceiver prefix to the GPT-2 decoder, restating
models/gpt2_model.py:178-196 in plain torch fp32.  Pinned by tests/golden/gpt2_prefix.pt (scripts/make_golden_pipeline.py:
the input of `transformer.drop` captured from the UNMODIFIED reference GPT, with the gradients its decoder sends back)."""
import torch


def embed_with_prefix(wte, wpe, idx, prefix=None):
    tok = wte[idx]                                              # :183 self.transformer.wte(idx)
    if prefix is not None:
        tok = torch.cat([prefix, tok], dim=1)                   # :187
    pos = torch.arange(0, tok.size(1), dtype=torch.long)        # :193
    return tok + wpe[pos]                                       # :196 (dropout p = 0)

"""Test infrastructure, not product code.  Generates tests/golden/umt5_tiny.pt: what the real third-party implementation
behind the reference's ``encode_prompt`` (delta_experiment/scripts/common.py:228-255 -> ``transformers.UMT5EncoderModel``,
imported from this image's site-packages, transformers 5.5) returns as ``last_hidden_state`` on the seeded tiny encoder
of ``oracle/umt5_oracle.py`` -- in fp32 (the parity reference) and with the whole model in bf16 (what the reference
runs: ``torch_dtype=dtype``, common.py:62-64).  Cases cover right-padded masks, a sequence that is not a multiple of the
kernel's 64-key block and more than one key block.

Run here:  python oracle/make_golden_umt5.py"""
import pathlib
import sys

import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle.umt5_oracle import TINY, tiny_inputs, tiny_state  # noqa: E402

CASES = [dict(batch=2, n_tok=96, seed=1), dict(batch=1, n_tok=200, seed=2)]


def hf_model(cfg, state, dtype):
    from transformers import UMT5Config, UMT5EncoderModel
    c = UMT5Config(vocab_size=cfg["vocab_size"], d_model=cfg["d_model"], d_kv=cfg["d_kv"], d_ff=cfg["d_ff"],
                   num_layers=cfg["num_layers"], num_heads=cfg["num_heads"],
                   relative_attention_num_buckets=cfg["relative_attention_num_buckets"],
                   relative_attention_max_distance=cfg["relative_attention_max_distance"],
                   layer_norm_epsilon=cfg["layer_norm_epsilon"], feed_forward_proj="gated-gelu", dropout_rate=0.0)
    m = UMT5EncoderModel(c).eval()
    missing, unexpected = m.load_state_dict(state, strict=False)
    assert not unexpected and all("embed_tokens" in k or "shared" in k for k in missing), (missing, unexpected)
    return m.to(dtype)


def main():
    import transformers
    state = tiny_state(TINY, seed=0)
    m32, m16 = hf_model(TINY, state, torch.float32), hf_model(TINY, state, torch.bfloat16)
    out = {"cfg": TINY, "state_seed": 0, "transformers": transformers.__version__, "cases": []}
    for c in CASES:
        ids, mask = tiny_inputs(TINY, **c)
        with torch.no_grad():
            y32 = m32(ids, mask).last_hidden_state
            y16 = m16(ids, mask).last_hidden_state
        out["cases"].append(dict(c, input_ids=ids, attention_mask=mask, last_hidden_state_fp32=y32,
                                 last_hidden_state_bf16=y16))
        print(c, float(y32.abs().mean()), float((y16.float() - y32).norm() / y32.norm()))
    torch.save(out, ROOT / "tests" / "golden" / "umt5_tiny.pt")


if __name__ == "__main__":
    main()

"""UMT5-xxl encoder (24 layers, d_model 4096, 64 heads x 64, d_ff 10240) at 512 tokens on one B200: this repo's
B200UMT5Encoder against transformers' UMT5EncoderModel in bf16 (the library path the reference runs, common.py:62-64,250)
-- parity of both against the fp32 transformers model on the same bf16-valued weights, then time per encode.
Writes one JSON object (argv[1], default gpurun_out/text_encoder_xxl.json).  Random-init weights, synthetic token ids."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from oracle.umt5_oracle import XXL  # noqa: E402  (config constants only; the checker here is transformers itself)


def gpu_state(cfg, seed=0):
    # transformers' own T5 initialisation scales (initializer_factor 1): q (d_model d_kv)^-1/2, k / v / wi d_model^-1/2,
    # o / wo fan_in^-1/2 -- scores of order one, so the bf16 library model itself stays close to fp32 over 24 layers
    g = torch.Generator(device="cuda").manual_seed(seed)
    d, inner, ff = cfg["d_model"], cfg["num_heads"] * cfg["d_kv"], cfg["d_ff"]

    def rn(*shape, std):
        return (torch.randn(*shape, generator=g, device="cuda") * std).to(torch.bfloat16)

    st = {"shared.weight": rn(cfg["vocab_size"], d, std=1.0)}
    st["encoder.embed_tokens.weight"] = st["shared.weight"]
    for l in range(cfg["num_layers"]):
        a, f = f"encoder.block.{l}.layer.0.", f"encoder.block.{l}.layer.1."
        st[a + "SelfAttention.q.weight"] = rn(inner, d, std=(d * cfg["d_kv"]) ** -0.5)
        st[a + "SelfAttention.k.weight"] = rn(inner, d, std=d ** -0.5)
        st[a + "SelfAttention.v.weight"] = rn(inner, d, std=d ** -0.5)
        st[a + "SelfAttention.o.weight"] = rn(d, inner, std=inner ** -0.5)
        st[a + "SelfAttention.relative_attention_bias.weight"] = rn(32, cfg["num_heads"], std=0.5)
        st[a + "layer_norm.weight"] = (1.0 + 0.1 * torch.randn(d, generator=g, device="cuda")).to(torch.bfloat16)
        st[f + "DenseReluDense.wi_0.weight"] = rn(ff, d, std=d ** -0.5)
        st[f + "DenseReluDense.wi_1.weight"] = rn(ff, d, std=d ** -0.5)
        st[f + "DenseReluDense.wo.weight"] = rn(d, ff, std=ff ** -0.5)
        st[f + "layer_norm.weight"] = (1.0 + 0.1 * torch.randn(d, generator=g, device="cuda")).to(torch.bfloat16)
    st["encoder.final_layer_norm.weight"] = (1.0 + 0.1 * torch.randn(d, generator=g, device="cuda")).to(torch.bfloat16)
    return st


def hf_model(cfg, state, dtype):
    from transformers import UMT5Config, UMT5EncoderModel
    c = UMT5Config(vocab_size=cfg["vocab_size"], d_model=cfg["d_model"], d_kv=cfg["d_kv"], d_ff=cfg["d_ff"],
                   num_layers=cfg["num_layers"], num_heads=cfg["num_heads"], relative_attention_num_buckets=32,
                   relative_attention_max_distance=128, layer_norm_epsilon=1e-6, feed_forward_proj="gated-gelu",
                   dropout_rate=0.0)
    with torch.device("cuda"):
        m = UMT5EncoderModel(c).to(dtype).eval()
    m.load_state_dict({k: v.to(dtype) for k, v in state.items()}, strict=False)
    return m


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/text_encoder_xxl.json"
    layers = int(sys.argv[2]) if len(sys.argv) > 2 else XXL["num_layers"]
    cfg = dict(XXL, num_layers=layers)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from longcat_video_tta_b200 import ops
    from longcat_video_tta_b200.text_encoder import B200UMT5Encoder
    B, N, n_valid = 1, 512, 77
    st = gpu_state(cfg)
    g = torch.Generator(device="cuda").manual_seed(1)
    ids = torch.randint(2, cfg["vocab_size"], (B, N), generator=g, device="cuda")
    mask = torch.zeros(B, N, dtype=torch.long, device="cuda")
    mask[:, :n_valid] = 1
    ids[mask == 0] = 0
    enc = B200UMT5Encoder(st, device="cuda", **{k: v for k, v in cfg.items() if k != "vocab_size"})
    res = {"config": {"workload": f"UMT5-xxl encoder, {layers} layers, d_model 4096, 64 heads x 64, d_ff 10240, "
                                  f"batch {B} x {N} tokens ({n_valid} real, right-padded), bf16, random init"}}
    ours = enc(ids, mask).last_hidden_state.float()
    with torch.no_grad():
        m16 = hf_model(cfg, st, torch.bfloat16)
        y16 = m16(ids, mask).last_hidden_state.float()
        m32 = hf_model(cfg, st, torch.float32)
        y32 = m32(ids, mask).last_hidden_state
        del m32
        torch.cuda.empty_cache()

    def rl2(a, b):
        return float((a - b).norm() / b.norm())

    def cs(a, b):
        return float(torch.dot(a.flatten(), b.flatten()) / (a.norm() * b.norm()))

    keep = mask.bool()
    res["parity"] = {"reference": "transformers UMT5EncoderModel fp32 (TF32 off) on the same bf16-valued weights",
                     "ours_rel_l2": rl2(ours, y32), "ours_cosine": cs(ours, y32),
                     "ours_rel_l2_real_tokens": rl2(ours[keep], y32[keep]), "ours_cosine_real_tokens": cs(ours[keep], y32[keep]),
                     "transformers_bf16_rel_l2": rl2(y16, y32), "transformers_bf16_cosine": cs(y16, y32),
                     "transformers_bf16_rel_l2_real_tokens": rl2(y16[keep], y32[keep])}
    n0, c0 = ops.kernel_launches(), ops.LAUNCHES
    enc(ids, mask)
    res["launches_per_encode"] = ops.kernel_launches() - n0
    ms_ours = timed(lambda: enc(ids, mask), 20)
    with torch.no_grad():
        ms_hf = timed(lambda: m16(ids, mask), 20)
    d, inner, ff, H = cfg["d_model"], cfg["num_heads"] * cfg["d_kv"], cfg["d_ff"], cfg["num_heads"]
    rows = B * N
    flop = layers * (2 * rows * (4 * inner * d + 3 * d * ff) + 4 * B * H * N * N * cfg["d_kv"])
    wbytes = layers * (4 * inner * d + 3 * d * ff) * 2
    res["ours"] = {"ms_per_encode": ms_ours, "tflops": flop / ms_ours / 1e9, "weight_stream_gb_s": wbytes / ms_ours / 1e6}
    res["transformers_bf16_eager"] = {"ms_per_encode": ms_hf, "tflops": flop / ms_hf / 1e9,
                                      "weight_stream_gb_s": wbytes / ms_hf / 1e6, "torch": torch.__version__}
    res["speedup"] = ms_hf / ms_ours
    res["algorithmic_gflop_per_encode"] = flop / 1e9
    res["weight_bytes_per_encode_gb"] = wbytes / 1e9
    # per-kernel-family device time of one encode (CUDA events around each ABI call)
    fam = {}
    orig = ops._call

    def probe(name, *args):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(name, *args)
        e1.record()
        fam.setdefault(name, []).append((e0, e1))

    ops._call = probe
    enc(ids, mask)
    torch.cuda.synchronize()
    ops._call = orig
    res["per_call_ms"] = {k: {"calls": len(v), "ms": sum(a.elapsed_time(b) for a, b in v)} for k, v in fam.items()}
    print(json.dumps(res, indent=1))
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()

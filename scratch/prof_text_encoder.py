"""two UMT5-xxl-width layers at 512 tokens, three encodes: the ncu target for t5_attn_kernel / t5_rmsnorm_kernel /
gemm_kernel<128> (scratch/gpu_call31.sh)"""
import sys

import torch

sys.path.insert(0, ".")
from oracle.umt5_oracle import XXL  # noqa: E402  (config constants)
from scratch.bench_text_encoder import gpu_state  # noqa: E402
from longcat_video_tta_b200.text_encoder import B200UMT5Encoder  # noqa: E402

cfg = dict(XXL, num_layers=2, vocab_size=4096)
enc = B200UMT5Encoder(gpu_state(cfg), device="cuda", **{k: v for k, v in cfg.items() if k != "vocab_size"})
ids = torch.randint(2, 4096, (1, 512), device="cuda")
mask = torch.zeros(1, 512, dtype=torch.long, device="cuda")
mask[:, :77] = 1
for _ in range(3):
    out = enc(ids, mask).last_hidden_state
torch.cuda.synchronize()
print(float(out.float().abs().mean()))

"""GPU parity tests of every C-ABI kernel against plain fp32 PyTorch references (floating-point kernels:
tolerance = bf16 storage, rtol 2e-2 / cosine >= 0.999 as BASELINE.json's north_star states)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

BF16, F32 = torch.bfloat16, torch.float32


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from longcat_video_tta_b200 import ops as o
    o.selfcheck()  # fails loudly on anything that is not a B200
    return o


def cos(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)


def close(got, ref, rtol=2e-2, atol=None, min_cos=0.999):
    got, ref = got.float(), ref.float()
    assert torch.isfinite(got).all(), "non-finite values"
    atol = atol if atol is not None else 2e-2 * ref.abs().max().item() + 1e-6
    c = cos(got, ref).item() if ref.abs().max() > 0 else 1.0
    err = (got - ref).abs().max().item()
    assert c >= min_cos, f"cosine {c} (max err {err})"
    assert torch.allclose(got, ref, rtol=rtol, atol=atol), f"max err {err} vs atol {atol}, cosine {c}"


def rnd(*shape, dtype=BF16, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(dtype)


# ---------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (200, 512, 192), (1024, 1536, 512), (333, 64, 512), (1560, 4096, 4096)])
@pytest.mark.parametrize("mn", [False, True])
def test_gemm_plain(ops, M, N, K, mn):
    a = rnd(M, K, scale=0.5, seed=1)
    b = rnd(K, N, scale=0.5, seed=2) if mn else rnd(N, K, scale=0.5, seed=2)
    d = torch.empty(M, N, dtype=BF16, device="cuda")
    ops.gemm(M, N, [(a, b, K, mn, None)], ops.epi(ops.EPI_STORE, d))
    ref = a.float() @ (b.float() if mn else b.float().t())
    close(d, ref)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (512, 512, 1000), (1536, 512, 1024), (64, 4096, 3120), (4096, 64, 1024),
                                   (12288, 4096, 4680)])
def test_gemm_weight_gradient_form(ops, M, N, K):
    """dW [out, in] = dY^T X with dY [tokens, out] and X [tokens, in] read in place (both operands token-major): the
    product full-model TTA needs for every linear (run_full_tta.py:95-219 lets autograd form it)."""
    dy = rnd(K, M, scale=0.5, seed=3)
    x = rnd(K, N, scale=0.5, seed=4)
    dw = torch.empty(M, N, dtype=BF16, device="cuda")
    ops.gemm(M, N, [(dy, x, K, True, None, True)], ops.epi(ops.EPI_STORE, dw))
    close(dw, dy.float().t() @ x.float())
    # fp32 output, strided token-major views (a column slice of a fused buffer, as qkv / kv gradients are)
    big = rnd(K, M + 64, scale=0.5, seed=5)
    dwf = torch.empty(M, N, dtype=F32, device="cuda")
    ops.gemm(M, N, [(big[:, 64:], x, K, True, None, True)], ops.epi(ops.EPI_STORE_F32, dwf))
    close(dwf, big[:, 64:].float().t() @ x.float())


def test_gemm_two_segments_lora_and_bias_f32_out(ops):
    M, N, K, r = 1000, 768, 512, 16
    x, w = rnd(M, K, seed=1), rnd(N, K, scale=0.05, seed=2)
    t, b = rnd(M, r, seed=3), rnd(N, r, scale=0.1, seed=4)
    bias = rnd(N, seed=5)
    d = torch.empty(M, N, dtype=F32, device="cuda")
    ops.gemm(M, N, [(x, w, K, False, None), (t, b, r, False, None)], ops.epi(ops.EPI_STORE_F32, d, bias=bias))
    ref = x.float() @ w.float().t() + t.float() @ b.float().t() + bias.float()
    close(d, ref)
    # backward form: dX = dY W + U A  (MN-major B operands)
    dy, u, a = rnd(M, N, seed=6), rnd(M, r, seed=7), rnd(r, K, scale=0.1, seed=8)
    dx = torch.empty(M, K, dtype=BF16, device="cuda")
    ops.gemm(M, K, [(dy, w, N, True, None), (u, a, r, True, None)], ops.epi(ops.EPI_STORE, dx))
    close(dx, dy.float() @ w.float() + u.float() @ a.float())


def test_gemm_epilogues(ops):
    M, N, K, tpf = 512, 512, 256, 128
    x, w, bias = rnd(M, K, seed=1), rnd(N, K, scale=0.1, seed=2), rnd(N, seed=3)
    acc = x.float() @ w.float().t() + bias.float()
    d = torch.empty(M, N, dtype=BF16, device="cuda")
    ops.gemm(M, N, [(x, w, K, False, None)], ops.epi(ops.EPI_GELU, d, bias=bias))
    close(d, torch.nn.functional.gelu(acc, approximate="tanh"))
    # gate + residual (+ branch copy)
    resid, gate = rnd(M, N, seed=4), rnd(M // tpf, 6 * N, dtype=F32, seed=5)
    g = gate[:, 2 * N:3 * N]
    d2 = torch.empty(M, N, dtype=BF16, device="cuda")
    ops.gemm(M, N, [(x, w, K, False, None)], ops.epi(ops.EPI_GATE_RESID, d, bias=bias, resid=resid, gate=g,
                                                     tokens_per_frame=tpf, d2=d2))
    close(d, resid.float() + g.repeat_interleave(tpf, 0) * acc)
    close(d2, acc)
    ops.gemm(M, N, [(x, w, K, False, None)], ops.epi(ops.EPI_GATE_RESID, d, bias=bias, resid=resid))
    close(d, resid.float() + acc)
    # SwiGLU with co-tiled w1 | w3
    F_ = 384
    w1, w3 = rnd(F_, K, scale=0.1, seed=6), rnd(F_, K, scale=0.1, seed=7)
    h = torch.empty(M, F_, dtype=BF16, device="cuda")
    h1 = torch.empty_like(h)
    h3 = torch.empty_like(h)
    ops.gemm(M, 2 * F_, [(x, w1, K, False, w3)], ops.epi(ops.EPI_SWIGLU, h, d2=h1, d3=h3))
    r1, r3 = x.float() @ w1.float().t(), x.float() @ w3.float().t()
    close(h1, r1)
    close(h3, r3)
    close(h, torch.nn.functional.silu(r1) * r3)
    # SwiGLU backward epilogue: dh = dy W2 ; (dh1, dh3)
    w2, dy = rnd(K, F_, scale=0.1, seed=8), rnd(M, K, seed=9)
    dh1 = torch.empty(M, F_, dtype=BF16, device="cuda")
    dh3 = torch.empty_like(dh1)
    ops.gemm(M, F_, [(dy, w2, K, True, None)], ops.epi(ops.EPI_SWIGLU_BWD, dh1, d2=dh3, aux1=h1, aux2=h3))
    dh = dy.float() @ w2.float()
    h1f, h3f = h1.float(), h3.float()
    sig = torch.sigmoid(h1f)
    close(dh1, dh * h3f * sig * (1 + h1f * (1 - sig)))
    close(dh3, dh * h1f * sig)


def test_lora_linear_fwd_bwd(ops):
    n, cin, cout, r, scale = 1560, 512, 1536, 16, 2.0
    x, w, bias = rnd(n, cin, seed=1), rnd(cout, cin, scale=0.05, seed=2), rnd(cout, seed=3)
    A, B = rnd(r, cin, scale=0.05, seed=4), rnd(cout, r, scale=0.05, seed=5)
    y = torch.empty(n, cout, dtype=BF16, device="cuda")
    xa = torch.empty(n, r, dtype=BF16, device="cuda")
    ops.lora_linear_fwd(x, w, ops.epi(ops.EPI_STORE, y, bias=bias), A=A, B=B, XA=xa, scale=scale)
    xf, wf, Af, Bf = x.float(), w.float(), A.float(), B.float()
    close(xa, scale * xf @ Af.t())
    close(y, xf @ wf.t() + bias.float() + scale * (xf @ Af.t()) @ Bf.t())
    dy = rnd(n, cout, seed=6)
    dx = torch.empty(n, cin, dtype=BF16, device="cuda")
    u = torch.empty(n, r, dtype=BF16, device="cuda")
    dA = torch.zeros(cin, r, dtype=F32, device="cuda")
    dB = torch.zeros(cout, r, dtype=F32, device="cuda")
    ops.lora_linear_bwd(dy, w, ops.epi(ops.EPI_STORE, dx), x=x, A=A, B=B, XA=xa, U=u, dA_acc=dA, dB_acc=dB, scale=scale)
    dyf = dy.float()
    close(dx, dyf @ wf + scale * (dyf @ Bf) @ Af)
    close(dB, scale * dyf.t() @ (xf @ Af.t()))
    close(dA.t(), scale * (dyf @ Bf).t() @ xf)


# ---------------------------------------------------------------------------------------------- elementwise
@pytest.mark.parametrize("C", [512, 4096])
def test_ln_mod_fwd_bwd(ops, C):
    rows, tpf = 3 * 100, 100
    x = rnd(rows, C, seed=1)
    mod = rnd(3, 6 * C, dtype=F32, scale=0.5, seed=2)
    shift, scale = mod[:, :C], mod[:, C:2 * C]
    y = torch.empty_like(x)
    ops.ln_mod_fwd(y, x, scale, shift, tokens_per_frame=tpf)
    xr = x.float().requires_grad_(True)
    sc, sh = scale.clone().requires_grad_(True), shift.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (C,), eps=1e-6) * (1 + sc.repeat_interleave(tpf, 0)) + sh.repeat_interleave(tpf, 0)
    close(y, ref)
    dy, dres = rnd(rows, C, seed=3), rnd(rows, C, seed=4)
    ref.backward(dy.float())
    dx = torch.empty_like(x)
    dsc = torch.zeros(3, C, dtype=F32, device="cuda")
    dsh = torch.zeros(3, C, dtype=F32, device="cuda")
    ops.ln_mod_bwd(dx, dy, x, scale, dx_resid=dres, tokens_per_frame=tpf, dscale_acc=dsc, dshift_acc=dsh)
    close(dx, xr.grad + dres.float())
    close(dsc, sc.grad)
    close(dsh, sh.grad)
    # affine form (pre_crs_attn_norm)
    wgt, b = (1 + 0.1 * rnd(C, dtype=F32, seed=5)).to(BF16), rnd(C, scale=0.1, seed=6)
    ops.ln_mod_fwd(y, x, wgt, b, tokens_per_frame=tpf, affine=True)
    xr2 = x.float().requires_grad_(True)
    w2, b2 = wgt.float().requires_grad_(True), b.float().requires_grad_(True)
    ref2 = torch.nn.functional.layer_norm(xr2, (C,), w2, b2, eps=1e-6)
    close(y, ref2)
    ref2.backward(dy.float())
    dw = torch.zeros(C, dtype=F32, device="cuda")
    db = torch.zeros(C, dtype=F32, device="cuda")
    ops.ln_mod_bwd(dx, dy, x, wgt, tokens_per_frame=tpf, affine=True, dscale_acc=dw, dshift_acc=db)
    close(dx, xr2.grad)
    close(dw, w2.grad)
    close(db, b2.grad)


def test_qk_rmsnorm_rope_matches_oracle(ops):
    from oracle.dit_oracle import RMSNormFP32, RotaryPositionalEmbedding3D
    T, Hh, Ww, H, D = 3, 5, 7, 4, 128
    n = T * Hh * Ww
    qkv = rnd(n, 3 * H * D, seed=1)
    qn, kn = RMSNormFP32(D).cuda(), RMSNormFP32(D).cuda()
    with torch.no_grad():
        qn.weight.copy_(1 + 0.1 * torch.randn(D, device="cuda"))
        kn.weight.copy_(1 + 0.1 * torch.randn(D, device="cuda"))
    rope = RotaryPositionalEmbedding3D(D)
    out = torch.empty(n, 2 * H * D, dtype=BF16, device="cuda")
    wq, wk = qn.weight.detach().to(BF16), kn.weight.detach().to(BF16)
    ops.qk_rmsnorm_rope_fwd(out, qkv, wq, wk, H, H, grid_hw=(Hh, Ww))
    x = qkv.float().view(n, 3, H, D).requires_grad_(True)
    q, k = x[:, 0].transpose(0, 1)[None], x[:, 1].transpose(0, 1)[None]  # [1,H,n,D]
    qn.weight.data, kn.weight.data = wq.float(), wk.float()
    qr, kr = rope(qn(q), kn(k), (T, Hh, Ww))
    ref = torch.cat([qr[0].transpose(0, 1).reshape(n, H * D), kr[0].transpose(0, 1).reshape(n, H * D)], dim=1)
    close(out, ref)
    dy = rnd(n, 2 * H * D, seed=2)
    ref.backward(dy.float())
    dx = torch.zeros(n, 3 * H * D, dtype=BF16, device="cuda")
    dwq = torch.zeros(D, dtype=F32, device="cuda")
    dwk = torch.zeros(D, dtype=F32, device="cuda")
    ops.qk_rmsnorm_rope_bwd(dx, dy, qkv, wq, wk, H, H, grid_hw=(Hh, Ww), dwq_acc=dwq, dwk_acc=dwk)
    close(dx[:, :2 * H * D], x.grad.reshape(n, 3 * H * D)[:, :2 * H * D])
    close(dwq, qn.weight.grad)
    close(dwk, kn.weight.grad)
    # offset rows (noise tokens only) and no-rope variant
    ops.qk_rmsnorm_rope_fwd(out[: n - 35], qkv[35:], wq, wk, H, H, grid_hw=(Hh, Ww), row_offset=35)
    close(out[: n - 35], ref[35:].detach())


def test_gate_mul_and_dgate(ops):
    rows, C, tpf = 256, 512, 64
    dx, br = rnd(rows, C, seed=1), rnd(rows, C, seed=2)
    gate = rnd(4, 6 * C, dtype=F32, seed=3)[:, 2 * C:3 * C]
    dy = torch.empty_like(dx)
    dg = torch.zeros(4, C, dtype=F32, device="cuda")
    ops.gate_mul(dy, dx, gate, tokens_per_frame=tpf, branch=br, dgate_acc=dg)
    close(dy, dx.float() * gate.repeat_interleave(tpf, 0))
    close(dg, (dx.float() * br.float()).view(4, tpf, C).sum(1))


def test_noise_patchify_unpatchify_mse(ops):
    from oracle.dit_oracle import build_oracle_dit
    from oracle import tta_oracle as T
    Tc, Tt, H, W = 2, 3, 8, 12
    cond, tgt, eps = rnd(1, 16, Tc, H, W, seed=1), rnd(1, 16, Tt, H, W, seed=2), rnd(1, 16, Tt, H, W, seed=3)
    sigma = torch.tensor([0.37], device="cuda")
    n = (Tc + Tt) * (H // 2) * (W // 2)
    nt = Tt * (H // 2) * (W // 2)
    P = torch.empty(n, 64, dtype=BF16, device="cuda")
    V = torch.empty(nt, 64, dtype=F32, device="cuda")
    ts = torch.empty(Tc + Tt, dtype=F32, device="cuda")
    ops.noise_patchify(P, V, ts, cond[0], tgt[0], eps[0], sigma)
    hidden, timestep, n_cond = T.build_step_inputs(cond, tgt, sigma, eps, BF16)
    assert torch.equal(ts, timestep[0].float())
    # P @ conv_weight^T must equal Conv3d patch embedding of `hidden`
    dit = build_oracle_dit("tiny", seed=0).cuda()
    wpe = dit.x_embedder.proj.weight.reshape(512, 64)
    ref = dit.x_embedder(hidden.float())[0] - dit.x_embedder.proj.bias
    close(P.float() @ wpe.t(), ref, rtol=1e-2)
    # velocity target / unpatchify / MSE
    vel = (eps - tgt).float()  # bf16 subtraction then upcast (common.py:486)
    lat = torch.empty(16, Tt, H, W, dtype=F32, device="cuda")
    ops.unpatchify(lat, V, Tt, H, W)
    assert torch.equal(lat, vel[0])
    tok = rnd(nt, 64, dtype=F32, seed=4)
    ops.unpatchify(lat, tok, Tt, H, W)
    assert torch.equal(lat, dit.unpatchify(tok[None], Tt, H // 2, W // 2)[0])
    loss = torch.zeros(1, dtype=F32, device="cuda")
    dpred = torch.empty(nt, 64, dtype=BF16, device="cuda")
    ops.mse_fwd_bwd(loss, dpred, tok, V)
    close(loss, torch.nn.functional.mse_loss(tok, V)[None], rtol=1e-4, atol=1e-6)
    close(dpred, 2 * (tok - V) / tok.numel())


def test_timestep_embedding_and_adaln(ops):
    from oracle.dit_oracle import build_oracle_dit
    dit = build_oracle_dit("tiny", seed=0).cuda()
    ts = torch.tensor([0.0, 0.0, 371.0, 371.0, 998.0], device="cuda")
    F_ = torch.empty(5, 256, dtype=F32, device="cuda")
    ops.timestep_sinusoid(F_, ts)
    close(F_, dit.t_embedder.timestep_embedding(ts, 256), rtol=1e-3, atol=2e-3)
    l0, l2 = dit.t_embedder.mlp[0], dit.t_embedder.mlp[2]
    h = torch.empty(5, 512, dtype=F32, device="cuda")
    t = torch.empty(5, 512, dtype=F32, device="cuda")
    ops.skinny_linear(h, F_, l0.weight.to(BF16), l0.bias.to(BF16))
    ops.skinny_linear(t, h, l2.weight.to(BF16), l2.bias.to(BF16), act=1)
    close(t, dit.t_embedder(ts), rtol=2e-2)
    ada = dit.blocks[0].adaLN_modulation[1]
    mod = torch.empty(5, 6 * 512, dtype=F32, device="cuda")
    ops.skinny_linear(mod, t, ada.weight, ada.bias, act=1)
    ref_t = t.clone().requires_grad_(True)
    ref = torch.nn.functional.linear(torch.nn.functional.silu(ref_t), ada.weight, ada.bias)
    close(mod, ref, rtol=1e-3)
    dmod = rnd(5, 6 * 512, dtype=F32, seed=3)
    ref.backward(dmod)
    dt = torch.empty(5, 512, dtype=F32, device="cuda")
    ops.skinny_linear_bwd(dt, dmod, t, ada.weight, act=1)
    close(dt, ref_t.grad, rtol=1e-3)


# ---------------------------------------------------------------------------------------------- optimizer
def test_clip_and_adamw_match_oracle(ops):
    from oracle import tta_oracle as T
    shapes = [(16, 512), (1536, 16), (512,), (16, 512)]
    params = [rnd(*s, dtype=F32, scale=0.05, seed=i) for i, s in enumerate(shapes)]
    grads = [rnd(*s, dtype=F32, scale=0.3, seed=10 + i) for i, s in enumerate(shapes)]
    for mode in ("fp32", "bf16_master", "bf16_faithful"):
        dev_p, entries = [], []
        for i, (p, g) in enumerate(zip(params, grads)):
            tr = i == 3  # transposed gradient layout for one LoRA "down" matrix
            gp = g.t().contiguous() if tr else g.clone()
            if mode == "fp32":
                pp = p.clone()
                e = dict(param=pp, grad=gp, exp_avg=torch.zeros_like(pp), exp_avg_sq=torch.zeros_like(pp))
            elif mode == "bf16_master":
                pp = p.to(BF16)
                e = dict(param=pp, master=pp.float(), grad=gp, exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p))
            else:
                pp = p.to(BF16)
                e = dict(param=pp, grad=gp, exp_avg=torch.zeros_like(pp), exp_avg_sq=torch.zeros_like(pp))
            e["grad_transposed"] = tr
            entries.append(e)
            dev_p.append(pp)
        tl = ops.TensorList(entries, "cuda")
        # oracle
        rp = [(p if mode == "fp32" else p.to(BF16).float()).clone() for p in params]
        rg = [g.clone() for g in grads]
        rm, rv = [torch.zeros_like(p) for p in rp], [torch.zeros_like(p) for p in rp]
        total = T.clip_grad_norm(rg, 1.0)
        for step in (1, 2, 3):
            tl.clip_coef(1.0)
            tl.adamw(lr=2e-4 * step / 3, eps=1e-8, weight_decay=0.01, step=step, faithful_bf16=(mode == "bf16_faithful"))
            T.adamw_step(rp, rg, rm, rv, step, 2e-4 * step / 3, eps=1e-8, wd=0.01)
        close(tl.total_norm, total[None], rtol=1e-4, atol=1e-6)
        if mode == "bf16_faithful":
            # compare with torch's own foreach AdamW on bf16 tensors (what the reference executes)
            tp = [torch.nn.Parameter(p.to(BF16)) for p in params]
            opt = torch.optim.AdamW(tp, lr=2e-4, betas=(0.9, 0.999), weight_decay=0.01, eps=1e-8)
            for step in (1, 2, 3):
                for q, g in zip(tp, grads):
                    q.grad = g.to(BF16)
                torch.nn.utils.clip_grad_norm_(tp, 1.0)
                for pg in opt.param_groups:
                    pg["lr"] = 2e-4 * step / 3
                opt.step()
            for got, want in zip(dev_p, tp):
                diff = (got.float() - want.detach().float()).abs()
                assert diff.max() <= 2 * torch.finfo(BF16).eps * want.detach().float().abs().max()
        else:
            for got, want in zip(entries, rp):
                ref = want
                val = got["master"] if mode == "bf16_master" else got["param"]
                assert torch.allclose(val.float(), ref, rtol=1e-5, atol=1e-7), (val.float() - ref).abs().max()


# ---------------------------------------------------------------------------------------------- attention
def _sdpa_ref(q, k, v, segs, scale):
    """q,k,v [n,H,D] fp32 -> o [n,H,D], lse [H,n]"""
    n, H, D = q.shape
    o = torch.zeros_like(q)
    lse = torch.zeros(H, n, device=q.device)
    for (a, b, kv) in segs:
        s = torch.einsum("qhd,khd->hqk", q[a:b], k[:kv]) * scale
        lse[:, a:b] = torch.logsumexp(s, dim=-1)
        o[a:b] = torch.einsum("hqk,khd->qhd", torch.softmax(s, dim=-1), v[:kv])
    return o, lse


@pytest.mark.parametrize("n,H,segs", [
    (256, 2, [(0, 256, 256)]),
    (1024, 4, [(0, 512, 512), (512, 1024, 1024)]),          # tiny config: cond | noise
    (1000, 3, [(0, 390, 390), (390, 1000, 1000)]),          # ragged everything
    (3120, 2, [(0, 1560, 1560), (1560, 3120, 3120)]),       # 480p frame size
])
def test_attn_fwd(ops, n, H, segs):
    D = 128
    qkv = rnd(n, 3, H, D, seed=1)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]  # strided views of one fused buffer
    o = torch.zeros(n, H, D, dtype=BF16, device="cuda")
    lse = torch.zeros(H, n, dtype=F32, device="cuda")
    ops.attn_fwd(q, k, v, o, lse, segs, D ** -0.5)
    ro, rl = _sdpa_ref(q.float(), k.float(), v.float(), segs, D ** -0.5)
    close(o, ro)
    close(lse, rl, rtol=1e-3, atol=1e-2)


def test_attn_fwd_growing_row_maximum(ops):
    """Keys whose magnitude jumps from one 128-row K/V block to the next: the running row maximum grows by far more
    than the 2^8 stale-maximum threshold at blocks 2, 4 and 7, so the forward kernel has to take its O / l rescale
    path (random N(0,1) inputs never do)."""
    n, H, D = 1024, 2, 128
    segs = [(0, 512, 512), (512, 1024, 1024)]
    qkv = rnd(n, 3, H, D, seed=5).float()
    f = torch.tensor([0.25, 1.0, 4.0, 0.5, 10.0, 1.0, 2.0, 20.0], device="cuda").repeat_interleave(128)
    qkv[:, 1] *= f[:, None, None]
    qkv = qkv.to(BF16)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    o = torch.zeros(n, H, D, dtype=BF16, device="cuda")
    lse = torch.zeros(H, n, dtype=F32, device="cuda")
    ops.attn_fwd(q, k, v, o, lse, segs, D ** -0.5)
    ro, rl = _sdpa_ref(q.float(), k.float(), v.float(), segs, D ** -0.5)
    close(o, ro)
    close(lse, rl, rtol=1e-3, atol=1e-2)


def test_attn_fwd_cross_attention_shape(ops):
    """queries = noise tokens, keys = 128 packed text tokens (kv [M, 2, H, D])"""
    n, M, H, D = 700, 128, 4, 128
    q = rnd(n, H, D, seed=1)
    kv = rnd(M, 2, H, D, seed=2)
    o = torch.zeros(n, H, D, dtype=BF16, device="cuda")
    lse = torch.zeros(H, n, dtype=F32, device="cuda")
    ops.attn_fwd(q, kv[:, 0], kv[:, 1], o, lse, [(0, n, M)], D ** -0.5)
    ro, rl = _sdpa_ref(q.float(), kv[:, 0].float(), kv[:, 1].float(), [(0, n, M)], D ** -0.5)
    close(o, ro)
    close(lse, rl, rtol=1e-3, atol=1e-2)


@pytest.fixture(params=["fused", "split"])
def bwd_mode(request, monkeypatch):
    """both backward implementations behind b200tta_attn_bwd*: the fused five-product kernel (default) and the
    dq + dkv pair (B200TTA_ATTN_BWD=split; also what the block-sparse path runs)"""
    monkeypatch.setenv("B200TTA_ATTN_BWD", request.param)
    return request.param


@pytest.mark.parametrize("n,H,segs", [
    (256, 2, [(0, 256, 256)]),
    (1024, 4, [(0, 512, 512), (512, 1024, 1024)]),
    (1000, 3, [(0, 390, 390), (390, 1000, 1000)]),
    (3120, 2, [(0, 1560, 1560), (1560, 3120, 3120)]),
    (1337, 2, [(0, 200, 200), (200, 1337, 1337)]),           # no boundary on a tile edge
])
def test_attn_bwd(ops, bwd_mode, n, H, segs):
    D = 128
    scale = D ** -0.5
    qkv = rnd(n, 3, H, D, seed=1)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    o = torch.zeros(n, H, D, dtype=BF16, device="cuda")
    lse = torch.zeros(H, n, dtype=F32, device="cuda")
    ops.attn_fwd(q, k, v, o, lse, segs, scale)
    do = rnd(n, H, D, seed=2)
    dqkv = torch.full((n, 3, H, D), float("nan"), dtype=BF16, device="cuda")
    delta = torch.empty(H, n, dtype=F32, device="cuda")
    ops.attn_bwd(dqkv[:, 0], dqkv[:, 1], dqkv[:, 2], do, o, lse, delta, q, k, v, segs, scale)
    qf, kf, vf = (t.float().detach().requires_grad_(True) for t in (q, k, v))
    ro, _ = _sdpa_ref(qf, kf, vf, segs, scale)
    ro.backward(do.float())
    close(delta, (do.float() * o.float()).sum(-1).t(), rtol=2e-2)
    close(dqkv[:, 0], qf.grad)
    close(dqkv[:, 1], kf.grad)
    close(dqkv[:, 2], vf.grad)


def test_attn_bwd_cross_attention_shape(ops, bwd_mode):
    n, M, H, D = 700, 128, 4, 128
    scale = D ** -0.5
    q, kv = rnd(n, H, D, seed=1), rnd(M, 2, H, D, seed=2)
    o = torch.zeros(n, H, D, dtype=BF16, device="cuda")
    lse = torch.zeros(H, n, dtype=F32, device="cuda")
    segs = [(0, n, M)]
    ops.attn_fwd(q, kv[:, 0], kv[:, 1], o, lse, segs, scale)
    do = rnd(n, H, D, seed=3)
    dq = torch.empty_like(q)
    dkv = torch.empty_like(kv)
    delta = torch.empty(H, n, dtype=F32, device="cuda")
    ops.attn_bwd(dq, dkv[:, 0], dkv[:, 1], do, o, lse, delta, q, kv[:, 0], kv[:, 1], segs, scale)
    qf, kf, vf = (t.float().detach().requires_grad_(True) for t in (q, kv[:, 0], kv[:, 1]))
    ro, _ = _sdpa_ref(qf, kf, vf, segs, scale)
    ro.backward(do.float())
    close(dq, qf.grad)
    close(dkv[:, 0], kf.grad)
    close(dkv[:, 1], vf.grad)


# ---------------------------------------------------------------------------------------------- headline sizes
# BASELINE.json configs[1]: N = 37 440 tokens (6 240 context + 31 200 noised), 32 heads x 128.  A dense fp32 reference of
# the whole problem does not fit, so these tests use size-independent properties plus fp32 spot checks on slices.
HEAD_N, HEAD_NC, HEAD_H, HEAD_D = 37440, 6240, 32, 128
HEAD_SEGS = [(0, HEAD_NC, HEAD_NC), (HEAD_NC, HEAD_N, HEAD_N)]


@pytest.fixture(scope="module")
def headline_attention(ops):
    qkv = rnd(HEAD_N, 3, HEAD_H, HEAD_D, seed=11)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    o = torch.zeros(HEAD_N, HEAD_H, HEAD_D, dtype=BF16, device="cuda")
    lse = torch.zeros(HEAD_H, HEAD_N, dtype=F32, device="cuda")
    ops.attn_fwd(q, k, v, o, lse, HEAD_SEGS, HEAD_D ** -0.5)
    return q, k, v, o, lse


def test_attn_fwd_headline_size_properties_and_slices(ops, headline_attention):
    q, k, v, o, lse = headline_attention
    scale = HEAD_D ** -0.5
    # (1) softmax rows sum to one: with V = 1 the output is 1 whatever Q and K are
    ones = torch.ones_like(v)
    o1 = torch.empty_like(o)
    lse1 = torch.empty_like(lse)
    ops.attn_fwd(q, k, ones, o1, lse1, HEAD_SEGS, scale)
    assert (o1.float() - 1.0).abs().max().item() <= 2 ** -7
    assert torch.equal(lse1, lse)                      # the statistics do not depend on V
    # (2) fp32 reference on slices: first / last rows of each segment, two heads
    for (a, b, kv) in HEAD_SEGS:
        for rows in (slice(a, a + 192), slice(b - 200, b)):
            for h in (0, HEAD_H - 1):
                s = (q[rows, h].float() @ k[:kv, h].float().t()) * scale
                ref = torch.softmax(s, -1) @ v[:kv, h].float()
                close(o[rows, h], ref)
                close(lse[h, rows], torch.logsumexp(s, -1), rtol=1e-3, atol=1e-2)


def test_attn_bwd_headline_size_properties_and_slices(ops, bwd_mode, headline_attention):
    q, k, v, o, lse = headline_attention
    scale = HEAD_D ** -0.5
    do = rnd(HEAD_N, HEAD_H, HEAD_D, seed=12)
    dqkv = torch.full((HEAD_N, 3, HEAD_H, HEAD_D), float("nan"), dtype=BF16, device="cuda")
    delta = torch.empty(HEAD_H, HEAD_N, dtype=F32, device="cuda")
    ops.attn_bwd(dqkv[:, 0], dqkv[:, 1], dqkv[:, 2], do, o, lse, delta, q, k, v, HEAD_SEGS, scale)
    dq, dk, dv = dqkv[:, 0], dqkv[:, 1], dqkv[:, 2]
    assert torch.isfinite(dqkv.float()).all()
    # (1) checksum: every row of P sums to one, so sum_kv dV[kv] = sum_q dO[q] per head (fp32 accumulation of bf16 P)
    got, want = dv.float().sum(0), do.float().sum(0)
    assert ((got - want).abs().max() / want.abs().max()).item() < 2e-2
    # (2) linearity in dO (x2 is exact in bf16; only the summation order inside the kernels may differ run to run)
    dqkv2 = torch.empty_like(dqkv)
    ops.attn_bwd(dqkv2[:, 0], dqkv2[:, 1], dqkv2[:, 2], (do.float() * 2).to(BF16), o, lse, delta, q, k, v, HEAD_SEGS, scale)
    for a, b in ((dqkv2[:, 0], dq), (dqkv2[:, 1], dk), (dqkv2[:, 2], dv)):
        rel = ((a.float() - 2 * b.float()).norm() / (2 * b.float().norm())).item()
        assert rel < 5e-3, rel
    # (3) fp32 autograd on one head: dQ on row slices of both segments, dK / dV everywhere (needs every query)
    h = 5
    qf, kf, vf = (t[:, h].float().detach().requires_grad_(True) for t in (q, k, v))
    out = torch.zeros(HEAD_N, HEAD_D, device="cuda")
    chunks = []
    for (a, b, kv) in HEAD_SEGS:
        for r0 in range(a, b, 4096):                 # chunked over queries: a [4096, 37440] fp32 score block at a time
            r1 = min(b, r0 + 4096)
            s = (qf[r0:r1] @ kf[:kv].t()) * scale
            chunks.append(((torch.softmax(s, -1) @ vf[:kv]) * do[r0:r1, h].float()).sum())
    torch.stack(chunks).sum().backward()
    close(dq[:, h], qf.grad)
    close(dk[:, h], kf.grad)
    close(dv[:, h], vf.grad)


def test_gemm_headline_size_slices(ops):
    """qkv projection of the headline step: [37440, 4096] x [12288, 4096]^T, checked on random rows against fp32."""
    M, N, K = HEAD_N, 3 * 4096, 4096
    a, w = rnd(M, K, seed=21), rnd(N, K, scale=0.02, seed=22)
    bias = rnd(N, seed=23)
    out = torch.empty(M, N, dtype=BF16, device="cuda")
    ops.gemm(M, N, [(a, w, K, False, None)], ops.epi(ops.EPI_STORE, out, bias=bias))
    rows = torch.randint(0, M, (256,), device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    rows = torch.cat([rows, torch.tensor([0, 127, 128, M - 129, M - 128, M - 1], device="cuda")])
    ref = a[rows].float() @ w.float().t() + bias.float()
    close(out[rows], ref)
    # dX form (MN-major B, no transposed copy): [37440, 12288] x [12288, 4096]
    dx = torch.empty(M, K, dtype=BF16, device="cuda")
    ops.gemm(M, K, [(out, w, N, True, None)], ops.epi(ops.EPI_STORE, dx))
    close(dx[rows], out[rows].float() @ w.float())


# ---------------------------------------------------------------------------------------------- full-model TTA helpers
@pytest.mark.parametrize("rows,C", [(1000, 64), (3120, 4096), (513, 12288), (37, 40)])
def test_colsum(ops, rows, C):
    big = rnd(rows, C + 8, seed=31)
    a = big[:, 8:] if C % 8 == 0 else big[:, :C].contiguous()          # strided view (a slice of a fused buffer)
    out = torch.full((C,), float("nan"), dtype=F32, device="cuda")
    ops.colsum(out, a)
    close(out, a.float().sum(0), rtol=1e-3, atol=1e-2 * (rows ** 0.5))


def test_mt_sgd_matches_torch_sgd(ops):
    """b200tta_mt_sgd == torch.optim.SGD(momentum=0, weight_decay) on bf16 parameters with fp32 gradients: the vectorised
    path (numel % 8 == 0), the scalar path, an fp32 parameter, a clip coefficient and a gradient scale."""
    shapes = [(64, 128), (7, 9), (4096,)]
    g = torch.Generator(device="cuda").manual_seed(7)
    params = [torch.randn(s, generator=g, device="cuda").to(BF16 if i < 2 else F32) for i, s in enumerate(shapes)]
    grads = [torch.randn(s, generator=g, device="cuda") for s in shapes]
    ref = [p.clone().float() for p in params]
    entries = [dict(param=p, grad=gr) for p, gr in zip(params, grads)]
    tl = ops.TensorList(entries, torch.device("cuda"))
    lr, wd, gs = 0.5, 0.01, 0.25
    tl.clip_coef(1.0, grad_scale=gs)
    total = torch.sqrt(sum((gr.float() ** 2).sum() for gr in grads)) * gs
    coef = min(1.0, 1.0 / (total.item() + 1e-6))
    tl.sgd(lr=lr, weight_decay=wd, grad_scale=gs)
    for p, r, gr in zip(params, ref, grads):
        want = r - lr * (gr * gs * coef + wd * r)
        if p.dtype == BF16:
            want = want.to(BF16)
        assert torch.equal(p, want.to(p.dtype)) or (p.float() - want.float()).abs().max() <= 2 ** -8 * want.float().abs().max()


def test_gather_rows_roundtrip_at_720p_size(ops):
    """block-sparse path: token order -> block-major -> token order is the identity (92 160 tokens x 4096, strided views)"""
    from longcat_video_tta_b200 import bsa
    n, H, D = 92160, 32, 128
    perm, inv = bsa.block_permutation(24, 48, 80, (4, 4, 8), device="cuda")
    qk = rnd(n, 2 * H * D, seed=41).view(n, 2 * H, D)
    q = qk[:, :H]                                            # row stride 2 C, as in the engine
    b = torch.empty(n, H, D, dtype=BF16, device="cuda")
    ops.gather_rows(b, q, perm)
    assert torch.equal(b, q.index_select(0, perm))
    back = torch.full_like(qk, float("nan"))
    ops.gather_rows(back[:, :H], b, inv)
    assert torch.equal(back[:, :H], q)
    assert torch.isnan(back[:, H:].float()).all()            # nothing written outside the destination view


def test_weight_gradient_gemm_headline_size_slices(ops):
    """dW of the qkv projection at the headline token count: [12288, 4096] = dY[37440, 12288]^T X[37440, 4096], fp32 out,
    checked on random rows against fp32 and through linearity (dW(2 dY) = 2 dW(dY) exactly: x2 is exact in bf16)."""
    K, M, N = HEAD_N, 3 * 4096, 4096
    dy, x = rnd(K, M, scale=0.05, seed=51), rnd(K, N, scale=0.5, seed=52)
    dw = torch.empty(M, N, dtype=F32, device="cuda")
    ops.gemm(M, N, [(dy, x, K, True, None, True)], ops.epi(ops.EPI_STORE_F32, dw))
    rows = torch.randint(0, M, (96,), device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    ref = dy[:, rows].float().t() @ x.float()
    close(dw[rows], ref)
    dw2 = torch.empty_like(dw)
    ops.gemm(M, N, [((dy.float() * 2).to(BF16), x, K, True, None, True)], ops.epi(ops.EPI_STORE_F32, dw2))
    assert ((dw2 - 2 * dw).norm() / (2 * dw).norm()).item() < 1e-5


import sys, torch
sys.path.insert(0, '.')
import torch.nn.functional as F
from oracle.dit_oracle import build_oracle_dit, modulate_fp32, _sdpa
from oracle.make_golden import tiny_inputs, tiny_split
from oracle import tta_oracle as T
from longcat_video_tta_b200.dit import B200DiT
BF16=torch.bfloat16
def cmp(name, a, o):
    a, o = a.float(), o.float()
    c = (a.flatten()@o.flatten()/(a.norm()*o.norm()+1e-30)).item()
    print(f"{name:10s} cos {c:.6f} rel {((a-o).norm()/o.norm()).item():.5f} max|ref| {o.abs().max().item():.3f}")
latents, prompt, mask = tiny_inputs(); cond, train, _ = tiny_split(latents)
oracle = build_oracle_dit("tiny", seed=0)
with torch.no_grad():
    for p in oracle.parameters(): p.copy_(p.to(BF16).float())
dit = B200DiT.from_oracle(oracle); oracle = oracle.cuda()
torch.manual_seed(42); sigma = torch.rand(1)*0.999+0.001; eps = torch.randn_like(train).to(BF16)
c_,t_,p_ = cond.to(BF16).cuda(), train.to(BF16).cuda(), prompt.to(BF16).cuda()
hidden, timestep, n_cond = T.build_step_inputs(c_, t_, sigma.cuda(), eps.cuda(), BF16)
dit.engine.debug = {}
with torch.no_grad():
    got = dit(hidden, timestep, p_, mask.cuda(), num_cond_latents=n_cond)
    dbg = dit.engine.debug
    ws = dit.engine.ws
    tt = oracle.t_embedder(timestep.float().flatten()).reshape(1, 4, -1)
    cmp("t", ws.t, tt[0])
    y, ysl = oracle.embed_text(p_.float(), mask.cuda())
    cmp("y", ws.y, y[0])
    cmp("x0", ws.xs[0], oracle.x_embedder(hidden.float())[0])
    Tn, tpf, C, H, Nc = 4, 256, 512, 4, 512
    for b, blk in enumerate(oracle.blocks):
        x = dbg[(b, "f_xin")][None]          # my block input, fp32
        mod = F.linear(F.silu(tt), blk.adaLN_modulation[1].weight, blk.adaLN_modulation[1].bias).unsqueeze(2)
        cmp(f"b{b} mod", dbg[(b,"f_mod")], mod[0,:,0])
        sh1, sc1, g1, sh2, sc2, g2 = mod.chunk(6, dim=-1)
        xm1 = modulate_fp32(blk.mod_norm_attn, x.view(1,Tn,-1,C), sh1, sc1).view(1,-1,C)
        cmp(f"b{b} xm1", dbg[(b,"f_xm1")], xm1[0])
        qkv = blk.attn.qkv(xm1); cmp(f"b{b} qkv", dbg[(b,"f_qkv")], qkv[0])
        q,k,v = qkv.view(1,-1,3,H,128).permute(2,0,3,1,4).unbind(0)
        q,k = blk.attn.q_norm(q), blk.attn.k_norm(k); q,k = blk.attn.rope_3d(q,k,(4,16,16))
        mine_qk = dbg[(b,"f_qk")].view(-1,2,H,128)
        cmp(f"b{b} q", mine_qk[:,0], q[0].transpose(0,1)); cmp(f"b{b} k", mine_qk[:,1], k[0].transpose(0,1))
        oc = _sdpa(q[:,:,:Nc],k[:,:,:Nc],v[:,:,:Nc]); on = _sdpa(q[:,:,Nc:],k,v)
        o = torch.cat([oc,on],2).transpose(1,2).reshape(1,-1,C); cmp(f"b{b} o", dbg[(b,"f_o")], o[0])
        x1 = x + (g1 * blk.attn.proj(o).view(1,Tn,-1,C)).view(1,-1,C); cmp(f"b{b} x1", dbg[(b,"f_x1")], x1[0])
        xn = blk.pre_crs_attn_norm(x1); cmp(f"b{b} xn", dbg[(b,"f_xn")], xn[0,Nc:])
        ca = blk.cross_attn
        qc = ca.q_linear(xn[:,Nc:]); cmp(f"b{b} qc", dbg[(b,"f_qc")], qc[0])
        kv = ca.kv_linear(y); cmp(f"b{b} kvc", dbg[(b,"f_kvc")], kv[0])
        cro = ca(xn, y, ysl, num_cond_latents=2, shape=(4,16,16))
        x2 = x1 + cro; cmp(f"b{b} x2", dbg[(b,"f_x2")], x2[0])
        xm2 = modulate_fp32(blk.mod_norm_ffn, x2.view(1,Tn,-1,C), sh2, sc2).view(1,-1,C); cmp(f"b{b} xm2", dbg[(b,"f_xm2")], xm2[0])
        h1, h3 = blk.ffn.w1(xm2), blk.ffn.w3(xm2); cmp(f"b{b} h1", dbg[(b,"f_h1")], h1[0]); cmp(f"b{b} h3", dbg[(b,"f_h3")], h3[0])
        h = F.silu(h1)*h3; cmp(f"b{b} h", dbg[(b,"f_h")], h[0])
        xo = x2 + (g2 * blk.ffn.w2(h).view(1,Tn,-1,C)).view(1,-1,C); cmp(f"b{b} xout", dbg[(b,"f_xout")], xo[0])

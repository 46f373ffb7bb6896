import sys, copy, torch
sys.path.insert(0, '.')
from oracle.dit_oracle import build_oracle_dit
from oracle.make_golden import tiny_inputs, tiny_split
from oracle import tta_oracle as T
from longcat_video_tta_b200 import lora
from longcat_video_tta_b200.dit import B200DiT
from longcat_video_tta_b200.stepper import TTAStepper
BF16=torch.bfloat16
def cos(a,b):
    a,b=a.float().flatten(),b.float().flatten(); return (a@b/(a.norm()*b.norm()+1e-30)).item()
latents, prompt, mask = tiny_inputs(); cond, train, _ = tiny_split(latents)
oracle = build_oracle_dit("tiny", seed=0)
with torch.no_grad():
    for p in oracle.parameters(): p.copy_(p.to(BF16).float())
dit = B200DiT.from_oracle(oracle)
ffn = '--ffn' in sys.argv
torch.manual_seed(7); mods = lora.inject_lora_into_dit(dit, rank=16, alpha=32.0, target_ffn=ffn)
oracle = oracle.cuda()
torch.manual_seed(7); omods = T.inject_lora(oracle, rank=16, alpha=32.0, target_ffn=ffn)
# make B nonzero so that dA is exercised too
g = torch.Generator().manual_seed(3)
with torch.no_grad():
    for m, om in zip(mods, omods):
        b = (torch.randn(m.lora_up.weight.shape, generator=g) * 0.02).to(BF16)
        m.lora_up.weight.copy_(b.cuda()); om.lora_up.weight.copy_(b.float().cuda())
        om.lora_down.weight.copy_(m.lora_down.weight.float())
torch.manual_seed(42); sigma = torch.rand(1)*0.999+0.001; eps = torch.randn_like(train).to(BF16)
c_,t_,p_ = cond.to(BF16).cuda(), train.to(BF16).cuda(), prompt.to(BF16).cuda()
oparams = T.lora_parameters(omods)
for p in oparams: p.requires_grad_(True)
cap = {}
def mk(b, name, which):
    def hook(mod, gin, gout):
        cap[(b, name + "_gout")] = gout[0].detach()
        if gin[0] is not None: cap[(b, name + "_gin")] = gin[0].detach()
    return hook
for b, blk in enumerate(oracle.blocks):
    blk.register_full_backward_hook(mk(b, "blk", 0))
    blk.ffn.register_full_backward_hook(mk(b, "ffn", 0))
    blk.attn.register_full_backward_hook(mk(b, "attn", 0))
    blk.cross_attn.register_full_backward_hook(mk(b, "cross", 0))
oloss = T.fm_loss_given(oracle, c_.float(), t_.float(), p_.float(), mask.cuda(), sigma.cuda(), eps.float().cuda(), torch.float32)
ograds = torch.autograd.grad(oloss, oparams)
st = TTAStepper(dit)
dit.engine.debug = {}
loss = st.forward_backward(c_, t_, p_, mask.cuda(), sigma.cuda(), eps.cuda())
print("loss", loss.item(), "oracle", oloss.item())
names = [s.name for s in dit.engine.lora_sites()]
for i,s in enumerate(dit.engine.lora_sites()):
    gA, gB = s.param_grads()
    print(f"{s.name:36s} dA cos {cos(gA, ograds[2*i]):.5f} ratio {(gA.norm()/ograds[2*i].norm()).item():.4f} | dB cos {cos(gB, ograds[2*i+1]):.5f} ratio {(gB.norm()/ograds[2*i+1].norm()).item():.4f}")

dbg = dit.engine.debug
Nc = 512
mod = dit.engine.ws.mod.clone()   # block 0's after the backward (recomputed last)
import torch.nn.functional as F
C=512
for b in (1,0):
    blk = oracle.blocks[b]
    tt = oracle.t_embedder(torch.cat([torch.zeros(2), (sigma*1000).to(BF16).float().expand(2)]).cuda()).reshape(1,4,-1)
    omod = F.linear(F.silu(tt), blk.adaLN_modulation[1].weight, blk.adaLN_modulation[1].bias)[0]
    gate = omod[:, 5*C:6*C].repeat_interleave(256, 0)
    print(b, "oracle: ffn_gout vs gate*blk_gout", cos(cap[(b,"ffn_gout")][0], gate*cap[(b,"blk_gout")][0]))
    print(b, "mine  : ffn_gout vs gate*dx_out  ", cos(dbg[(b,"ffn_gout")], gate*dbg[(b,"dx_out")]))
    print(b, "mine ffn_gout vs oracle gate*blk_gout", cos(dbg[(b,"ffn_gout")], gate*cap[(b,"blk_gout")][0]))
    print(b, "per-channel: cos over rows of dx_out vs blk_gout, worst channels:")
    a, o = dbg[(b,"dx_out")][Nc:], cap[(b,"blk_gout")][0][Nc:]
    cc = (a*o).sum(0)/(a.norm(dim=0)*o.norm(dim=0)+1e-30)
    print("   min/mean channel cos", cc.min().item(), cc.mean().item(), " channel norm ratio range", (a.norm(dim=0)/o.norm(dim=0)).min().item(), (a.norm(dim=0)/o.norm(dim=0)).max().item())

pairs = [("dx_out","blk_gout"),("ffn_gout","ffn_gout"),("ffn_gin","ffn_gin"),("attn_gout","attn_gout"),("attn_gin","attn_gin"),("dx_in","blk_gin")]
for b in (1,0):
    for mine, theirs in pairs:
        if (b,theirs) not in cap: continue
        a = dbg[(b,mine)]; o = cap[(b,theirs)][0]
        print(f"block {b} {mine:14s} cos {cos(a,o):.5f} ratio {(a.norm()/o.norm()).item():.4f}   cond-rows cos {cos(a[:Nc],o[:Nc]):.5f} noise-rows cos {cos(a[Nc:],o[Nc:]):.5f}")
    if (b,"cross_gin") not in cap: continue
    a = dbg[(b,"cross_gin")]; o = cap[(b,"cross_gin")][0][Nc:]
    print(f"block {b} cross_gin      cos {cos(a,o):.5f} ratio {(a.norm()/o.norm()).item():.4f}")
    a = dbg[(b,"dx_after_ffn")][Nc:]; o = cap[(b,"cross_gout")][0][Nc:]
    print(f"block {b} cross_gout     cos {cos(a,o):.5f} ratio {(a.norm()/o.norm()).item():.4f}")

# --- direct check of the proj-site LoRA gradients of block 0 from the engine's own buffers
eng = dit.engine
s0 = eng.sites[0]["proj"]
g1 = dbg[(0, "attn_gout")]                       # dY of proj (fp32 copy of bf16)
xa = eng.ws.xa["proj"].view(-1)[: 1024 * 16].view(1024, 16).float()
dB_ref = g1.t() @ xa
print("proj dB: kernel vs torch(g1^T XA)", cos(s0.dB_acc, dB_ref), (s0.dB_acc.norm()/dB_ref.norm()).item())
o_ = eng.ws.o.float()
u_ref = s0.scale * (g1 @ s0.B.float())
dA_ref = (o_.t() @ u_ref)
print("proj dA^T: kernel vs torch", cos(s0.dA_acc, dA_ref), (s0.dA_acc.norm()/dA_ref.norm()).item())
print("XA proj vs torch", cos(xa, s0.scale * o_ @ s0.A.float().t()))
sq = eng.sites[0]["qkv"]
xaq = eng.ws.xa["qkv"].view(-1)[: 1024 * 16].view(1024, 16).float()
print("XA qkv vs torch", cos(xaq, sq.scale * eng.ws.xm1.float() @ sq.A.float().t()))
dqkv = eng.ws.dqkv.float()
print("qkv dB: kernel vs torch", cos(sq.dB_acc, dqkv.t() @ xaq))
print("oracle dB_proj vs torch-from-engine-buffers", cos(ograds[3], dB_ref))

# --- what plain bf16 PyTorch (the reference's arithmetic) achieves against fp32 on the same problem
import copy
ob = copy.deepcopy(oracle).to(BF16)
bmods = []
for blk in ob.blocks:
    bmods += [blk.attn.qkv, blk.attn.proj, blk.cross_attn.q_linear, blk.cross_attn.kv_linear, blk.cross_attn.proj]
bparams = T.lora_parameters(bmods)
for p in bparams: p.requires_grad_(True)
bloss = T.fm_loss_given(ob, c_, t_, p_, mask.cuda(), sigma.cuda(), eps.cuda(), BF16)
bgrads = torch.autograd.grad(bloss, bparams)
print("bf16 torch loss", bloss.item())
for i, s in enumerate(dit.engine.lora_sites()):
    gA, gB = s.param_grads()
    print(f"{s.name:36s} bf16-torch vs fp32: dA {cos(bgrads[2*i], ograds[2*i]):.5f} dB {cos(bgrads[2*i+1], ograds[2*i+1]):.5f} | mine vs fp32: dA {cos(gA, ograds[2*i]):.5f} dB {cos(gB, ograds[2*i+1]):.5f} | mine vs bf16-torch: dB {cos(gB, bgrads[2*i+1]):.5f}")

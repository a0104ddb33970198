"""GPU diagnostic: per-parameter fp32 gradient error table on the realistic workload."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.nn.functional as F
import torch_vae_b200 as tv
from oracle import vae_oracle as vo
torch.set_num_threads(len(os.sched_getaffinity(0)))
B, H, W = 32, 128, 128
params = vo.xavier_params(1, 10, H, W, [32, 64, 128, 256], seed=0)
x, eps = vo.synthetic_batch(B, H, W, 10, seed=1234)
r64 = vo.step(params, x, eps, dtype=torch.float64)
r32 = vo.step(params, x, eps)
m = tv.VanillaVAE(1, 10, (H, W), compute_dtype=torch.float32)
m.load_state_dict({**m.state_dict(), **params}); m = m.cuda().train()
out = m(x.cuda(), eps=eps.cuda()); l = m.loss(out); l["loss"].backward(); torch.cuda.synchronize()
print("loss err", vo.rel_err(l["loss"], r64["loss"]))
for k, p in m.named_parameters():
    fl = vo.grad_floor(r64["grads"], k)
    e = vo.rel_err(p.grad, r64["grads"][k], fl); n = vo.rel_err(r32["grads"][k], r64["grads"][k], fl)
    print(f"{k:28s} |g|={float(r64['grads'][k].norm()):.3e} err={e:.2e} noise={n:.2e} ratio={e/max(n,1e-9):.1f}")
# BN backward unit check at scale, with a large mean/sigma ratio
ops = tv.ops
for C, npix, mean in [(64, 32768, 0.0), (64, 32768, 5.0), (32, 131072, 5.0), (256, 2048, 5.0)]:
    g = torch.Generator().manual_seed(0)
    y = torch.randn(npix, C, generator=g) + mean
    d = torch.randn(npix, C, generator=g) * 0.01 + 0.02
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    yr = y.double().t().reshape(1, C, npix, 1).requires_grad_(True)
    gr, br = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    a = F.leaky_relu(F.batch_norm(yr, None, None, gr, br, True, 0.1, 1e-5), 0.01)
    a.backward(d.double().t().reshape(1, C, npix, 1))
    yd = y.cuda().view(1, npix, 1, C).contiguous()
    stats = torch.stack([y.double().sum(0), (y.double() ** 2).sum(0)]).flatten().cuda()
    vec = ops.bn_finalize(stats, gamma.cuda(), beta.cuda(), None, None, npix, True)
    dy, dg, db, dbias = ops.bn_lrelu_backward(d.cuda().view(1, npix, 1, C).contiguous(), yd, vec, gamma.cuda(), beta.cuda(), True)
    ref_dy = yr.grad.reshape(C, npix).t()
    print(f"BN bwd C={C} npix={npix} mean={mean}: dy {vo.rel_err(dy.view(npix, C).cpu(), ref_dy):.2e} dgamma {vo.rel_err(dg.cpu(), gr.grad):.2e} dbeta {vo.rel_err(db.cpu(), br.grad):.2e}")

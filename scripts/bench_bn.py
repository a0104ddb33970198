"""Micro-benchmark of the BatchNorm / LeakyReLU passes at the layer shapes of the bench (B=512, bf16)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_vae_b200 as tv
ops = tv.ops
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return min(ts)
for n_pix, C in [(2097152, 32), (524288, 64), (131072, 128), (32768, 256)]:
    y = torch.randn(n_pix, 1, 1, C, device="cuda").bfloat16().view(n_pix // 64, 8, 8, C)
    d = torch.randn_like(y)
    gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    vec = torch.stack([torch.ones(C), torch.zeros(C), torch.zeros(C), torch.ones(C)]).cuda()
    nb = y.numel() * 2
    t1 = timeit(lambda: ops.bn_apply_lrelu(y, vec, beta))
    t2 = timeit(lambda: ops.bn_lrelu_backward(d, y, vec, gamma, beta, True))
    print(f"[{n_pix},{C}] apply {t1*1e3:6.1f} us ({2*nb/t1/1e6:5.0f} GB/s)   bwd reduce+apply {t2*1e3:6.1f} us ({5*nb/t2/1e6:5.0f} GB/s)")

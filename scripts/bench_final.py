"""Micro-benchmark of the fused output-block kernels at the bench shape (B=512, 128x128x32 bf16)."""
import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_vae_b200 as tv
ops = tv.ops
B = int(os.environ.get("B", "512")); H = W = 128; C = 32
iters = 5
y = torch.randn(B, H, W, C, device="cuda").bfloat16()
gamma = torch.ones(C, device="cuda"); beta = torch.zeros(C, device="cuda")
w_last = torch.randn(1, C, 3, 3, device="cuda") / math.sqrt(9 * C); b_last = torch.zeros(1, device="cuda")
dl = (torch.randn(B, 1, H, W, device="cuda") / (B * H * W)).bfloat16()
yy = y.double()
stats = torch.cat([yy.sum((0, 1, 2)), (yy * yy).sum((0, 1, 2))])
rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
vec = ops.bn_finalize(stats, gamma, beta, rm, rv, B * H * W, True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return min(ts)
nb = y.numel() * 2
for tc in (False, True):
    ops.set_final_tc(tc)
    t_f = timeit(lambda: ops.final_conv_fwd(y, vec, beta, w_last, b_last))
    t_b = timeit(lambda: ops.final_block_backward(dl, y, vec, gamma, beta, w_last, True))
    print(f"tc={tc}: fwd {t_f*1e3:7.1f} us ({(nb + B*H*W*6)/t_f/1e6:6.0f} GB/s)   bwd(reduce+apply) {t_b*1e3:7.1f} us ({(3*nb)/t_b/1e6:6.0f} GB/s)")

"""Timeline of CTA 0 of the streamed-weight tcgen05 conv (conv_umma_tma.cu) on one layer: where each pipeline
stage spends its cycles.   SHAPE=B,H,W,Ci,Co python scripts/trace_conv2.py <mode>"""
import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_vae_b200 as tv
ops = tv.ops
B, H, W, Ci, Co = [int(v) for v in os.environ.get('SHAPE', '512,32,32,64,128').split(',')]
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
x = torch.randn(B, H, W, Ci, device="cuda").bfloat16()
y = torch.randn(B, H // 2, W // 2, Co, device="cuda").bfloat16()
w = torch.randn(Co, Ci, 3, 3, device="cuda") / math.sqrt(9 * Ci)
wpk = ops.umma_tma_pack_weight(w, 2, mode)
inp = x if mode == 0 else y
run = lambda: ops.conv_umma_tma(inp, wpk, None, None, Ci, Co, 2, mode)
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
print(f"SHAPE {B},{H},{W},{Ci},{Co} mode {mode}: {e0.elapsed_time(e1) * 100:.1f} us per launch")
buf = torch.zeros(6 * 512, dtype=torch.int64, device="cuda")
ops._L().vae_umma_set_trace(buf.data_ptr())
run(); torch.cuda.synchronize()
ops._L().vae_umma_set_trace(None)
t = buf.view(6, 512).cpu()
n = int((t[0] > 0).sum()); nt = int((t[4] > 0).sum())
t0 = int(t[0][:n].min())
print(f"{n} stages, {nt} tiles traced; total span {int(t[3][:n].max()) - t0} cycles")
for s in range(min(n, 40)):
    print(f"stage {s:3d}: prod_begin {int(t[0][s])-t0:7d} prod_end {int(t[1][s])-t0:7d} full_seen {int(t[2][s])-t0:7d} mma_issued {int(t[3][s])-t0:7d}")
for i in range(min(nt, 10)):
    print(f"tile {i} (epilogue set {i % 2}): tfull_seen {int(t[4][i])-t0:7d} accumulator released {int(t[5][i])-t0:7d}")
if n > 12:
    a, b = 6, n - 2
    f = lambda v: float(v.float().mean())
    print("steady state means (cycles): between full_seen", f(t[2][a+1:b] - t[2][a:b-1]), "| producer wait for empty", f(t[0][a+1:b] - t[1][a:b-1]),
          "| producer issue", f(t[1][a:b] - t[0][a:b]), "| TMA issue_end->full_seen", f(t[2][a:b] - t[1][a:b]), "| mma issue", f(t[3][a:b] - t[2][a:b]))

"""Timeline of CTA 0 of the tcgen05 gather kernel (clock64 samples per role)."""
import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_vae_b200 as tv
ops = tv.ops
import os
B, H, W, Ci, Co = [int(v) for v in os.environ.get('SHAPE', '512,128,128,32,32').split(',')]
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
x = torch.randn(B, H, W, Ci, device="cuda").bfloat16()
y = torch.randn(B, H // 2, W // 2, Co, device="cuda").bfloat16()
w = torch.randn(Co, Ci, 3, 3, device="cuda") / math.sqrt(9 * Ci)
tma = len(sys.argv) > 2 and sys.argv[2] == 'tma'
wpk = ops.umma_tma_pack_weight(w, 2, mode) if tma else ops.umma_pack_weight(w, 2, mode)
run = (lambda: ops.conv_umma_tma(inp, wpk, None, None, Ci, Co, 2, mode)) if tma else (lambda: ops.conv_umma(inp, wpk, None, None, Ci, Co, 2, mode))
inp = x if mode == 0 else y
run(); torch.cuda.synchronize()
buf = torch.zeros(6 * 512, dtype=torch.int64, device="cuda")
ops._L().vae_umma_set_trace(buf.data_ptr())
run(); torch.cuda.synchronize()
ops._L().vae_umma_set_trace(None)
t = buf.view(6, 512).cpu()
t0 = int(t[0][t[0] > 0].min())
names = ["prod_issue_begin", "prod_issue_end", "mma_full_seen", "mma_issued", "epi_tfull_seen", "epi_tempty"]
print("k-block timeline (cycles since first issue), mode", mode)
for s in list(range(0, 40)) + list(range(200, 230)):
    print(s, " ".join(f"{names[r][:14]}={int(t[r][s]) - t0:7d}" for r in range(4) if t[r][s] > 0))
print("tiles:")
for i in list(range(0, 12)) + list(range(50, 60)):
    print(i, " ".join(f"{names[r]}={int(t[r][i]) - t0:7d}" for r in (4, 5) if t[r][i] > 0))
d = (t[2][100:500] - t[2][99:499]).float()
print("mean cycles between consecutive full_seen (steady state):", float(d.mean()))
d2 = (t[3][100:500] - t[2][100:500]).float(); print("mean mma issue duration:", float(d2.mean()))
d3 = (t[1][100:500] - t[0][100:500]).float(); print("mean producer issue duration:", float(d3.mean()))
d4 = (t[2][100:500] - t[1][100:500]).float(); print("mean issue_end -> full_seen:", float(d4.mean()))

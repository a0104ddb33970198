"""Micro-benchmark of the conv kernels at the default model's layer shapes (B=512).
usage: python scripts/bench_conv.py [filter] [--iters N]"""
import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_vae_b200 as tv
ops = tv.ops
flt = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else ""
iters = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 5
B = int(os.environ.get("B", "512"))
LAYERS = [  # name, Hi, Wi, Ci, Co  (equivalent conv, stride 2)
    ("L128_32_32", 128, 128, 32, 32), ("L64_32_64", 64, 64, 32, 64), ("L32_64_128", 32, 32, 64, 128), ("L16_128_256", 16, 16, 128, 256)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return min(ts), sum(ts) / len(ts)
for name, H, W, Ci, Co in LAYERS:
    x = torch.randn(B, H, W, Ci, device="cuda").bfloat16()
    y = torch.randn(B, H // 2, W // 2, Co, device="cuda").bfloat16()
    w = torch.randn(Co, Ci, 3, 3, device="cuda") / math.sqrt(9 * Ci)
    bias_o, bias_i = torch.zeros(Co, device="cuda"), torch.zeros(Ci, device="cuda")
    w0, w1 = ops.umma_pack_weight(w, 2, 0), ops.umma_pack_weight(w, 2, 1)
    st_o = torch.empty(2 * Co, dtype=torch.float64, device="cuda"); st_i = torch.empty(2 * Ci, dtype=torch.float64, device="cuda")
    xb, yb = x.numel() * 2, y.numel() * 2
    vec = torch.stack([torch.ones(Ci), torch.zeros(Ci), torch.zeros(Ci), torch.ones(Ci)]).cuda(); beta = torch.zeros(Ci, device="cuda")
    t0, t1 = ops.umma_tma_pack_weight(w, 2, 0), ops.umma_tma_pack_weight(w, 2, 1)
    r0 = ops.umma_rw_pack_weight(w, 0) if ops.umma_rw_eligible(B, H, W, Ci, Co, 2, 0) else None
    r1 = ops.umma_rw_pack_weight(w, 1) if ops.umma_rw_eligible(B, H, W, Ci, Co, 2, 1) else None
    cases = {
        "rw_fprop": (lambda: ops.conv_umma_rw(x, r0, bias_o, st_o, Ci, Co, 2, 0), xb + yb) if r0 is not None else None,
        "rw_dgrad": (lambda: ops.conv_umma_rw(y, r1, bias_i, st_i, Ci, Co, 2, 1), xb + yb) if r1 is not None else None,
        "tma_fprop": (lambda: ops.conv_umma_tma(x, t0, bias_o, st_o, Ci, Co, 2, 0), xb + yb),
        "tma_dgrad": (lambda: ops.conv_umma_tma(y, t1, bias_i, st_i, Ci, Co, 2, 1), xb + yb),
        "fprop": (lambda: ops.conv_umma(x, w0, bias_o, st_o, Ci, Co, 2, 0), xb + yb),
        "fprop_fused": (lambda: ops.conv_umma(x, w0, bias_o, st_o, Ci, Co, 2, 0, pro=(vec, beta)), xb + yb),
        "dgrad": (lambda: ops.conv_umma(y, w1, bias_i, st_i, Ci, Co, 2, 1), xb + yb),
        "wgrad": (lambda: ops.conv_weight_grad(x, y, 2, Ci, Co, H, W), xb + yb),
    }
    for cname, case in cases.items():
        if case is None:
            continue
        fn, nbytes = case
        full = f"{name}.{cname}"
        if flt and flt not in full:
            continue
        tmin, tavg = timeit(fn)
        print(f"{full:28s} min {tmin*1e3:8.1f} us  avg {tavg*1e3:8.1f} us   {nbytes/1e6:7.1f} MB  {nbytes/tmin/1e6:7.0f} GB/s  roofline {nbytes/6.53e12*1e6:6.1f} us")

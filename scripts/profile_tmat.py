"""One launch of the fused-phase transposed conv (conv_umma_tmat.cu) at the wide / long shape (64 -> 64 channels,
1x128x1024 rolls, B = 64) and at the train shape (128 -> 64 @ 32x32 out, B = 512), with and without the statistics
epilogue, for `ncu --set full` (scripts/gpu_ncu_tmat.sh).  Prints event timings when run plain."""
import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_vae_b200 as tv
ops = tv.ops
dev = "cuda"
CASES = [("wide convT 64->64 @128x1024 B=64", 64, 64, 64, 128, 1024), ("train convT 128->64 @32x32 B=512", 512, 64, 128, 32, 32)]
for name, B, ci, co, hi, wi in CASES:   # ci = high-res (output) channels, co = low-res (input) channels
    y = torch.randn(B, hi // 2, wi // 2, co, device=dev).bfloat16()
    w = torch.randn(co, ci, 3, 3, device=dev) / math.sqrt(9 * ci)
    cw = ops.ConvWeights(w, torch.bfloat16, 2)
    cw.umma(1, tma=True)
    bias = torch.zeros(ci, device=dev)
    for stats in (True, False):
        st = torch.empty(2 * ci, dtype=torch.float64, device=dev) if stats else None
        for _ in range(int(os.environ.get("REPS", "3"))):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            ops.conv_data_grad(y, cw, bias, st)
            e.record()
            torch.cuda.synchronize()
        print(f"{name} stats={stats}: {s.elapsed_time(e) * 1e3:.1f} us")

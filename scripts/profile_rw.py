"""Event timings of the resident-weight tcgen05 conv (conv_umma_rw.cu) at the train shapes (B = 512, 1x128x128 rolls):
the ConvTranspose2d forwards (bias + statistics) and the data gradients of the two outer blocks."""
import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_vae_b200 as tv
ops = tv.ops
dev = "cuda"
B = int(os.environ.get("B", "512"))
CASES = [("convT 32->32 @128x128", 32, 32, 128, 128), ("convT 64->32 @64x64", 32, 64, 64, 64)]
for name, ci, co, hi, wi in CASES:   # ci = high-res channels, co = low-res channels
    x = torch.randn(B, hi, wi, ci, device=dev).bfloat16()
    y = torch.randn(B, hi // 2, wi // 2, co, device=dev).bfloat16()
    w = torch.randn(co, ci, 3, 3, device=dev) / math.sqrt(9 * ci)
    cw = ops.ConvWeights(w, torch.bfloat16, 2)
    bias_t = torch.zeros(ci, device=dev); bias_c = torch.zeros(co, device=dev)
    for what in ("convT fwd (bias+stats)", "convT as dgrad (plain)", "conv fwd (bias+stats)", "conv as dgrad-of-convT (plain)"):
        for _ in range(int(os.environ.get("REPS", "5"))):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            st_t = torch.empty(2 * ci, dtype=torch.float64, device=dev); st_c = torch.empty(2 * co, dtype=torch.float64, device=dev)
            torch.cuda.synchronize()
            s.record()
            if what.startswith("convT fwd"): ops.conv_data_grad(y, cw, bias_t, st_t)
            elif what.startswith("convT as"): ops.conv_data_grad(y, cw, None, None)
            elif what.startswith("conv fwd"): ops.conv_forward(x, cw, bias_c, st_c)
            else: ops.conv_forward(x, cw, None, None)
            e.record()
            torch.cuda.synchronize()
        print(f"{name} {what}: {s.elapsed_time(e) * 1e3:.1f} us")

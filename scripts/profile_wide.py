"""One launch of every tensor-core conv of the wide / long configuration (BASELINE configs[3]: 1x128x1024 rolls,
widths [64,128,256,512], B = 64 per GPU) in a fixed order, for `ncu --set full` (scripts/gpu_ncu_wide.sh).
Prints the launch order so that the ncu rows can be labelled."""
import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_vae_b200 as tv
ops = tv.ops
B = int(os.environ.get("B", "64"))
dev = "cuda"
LAYERS = [  # name, Ci (high-res side), Co (low-res side), Hi, Wi (high-res extents)
    ("enc.conv1/dec.convT2", 64, 128, 64, 512),
    ("enc.conv2/dec.convT1", 128, 256, 32, 256),
    ("enc.conv3/dec.convT0", 256, 512, 16, 128),
]
order = []
for rep in range(int(os.environ.get("REPS", "2"))):
    for name, ci, co, hi, wi in LAYERS:
        x = torch.randn(B, hi, wi, ci, device=dev).bfloat16()
        w = torch.randn(co, ci, 3, 3, device=dev) / math.sqrt(9 * ci)
        cw = ops.ConvWeights(w, torch.bfloat16, 2)
        cw.umma(0, tma=True); cw.umma(1, tma=True)
        st = torch.empty(2 * co, dtype=torch.float64, device=dev)
        y = ops.conv_forward(x, cw, torch.zeros(co, device=dev), st)       # Conv2d forward (mode 0)
        st1 = torch.empty(2 * ci, dtype=torch.float64, device=dev)
        ops.conv_data_grad(y, cw, torch.zeros(ci, device=dev), st1)        # ConvTranspose2d forward / data gradient (mode 1)
        ops.conv_weight_grad(x, y, 2, ci, co, hi, wi)                      # weight gradient
        if rep == 0:
            n0 = 2 if co > 256 else 1
            order += [f"{name} fprop {ci}->{co} @{hi}x{wi}" + (f" slice {k}" if n0 > 1 else "") for k in range(n0)]
            order += [f"{name} dgrad/convT {co}->{ci} @{hi}x{wi}", f"{name} wgrad {ci}x{co} @{hi}x{wi}"]
torch.cuda.synchronize()
print("ORDER: " + " | ".join(order))

import os, sys, time, torch
sys.path.insert(0, "/root/repo")
import torch_vae_b200 as tv
from oracle import vae_oracle as vo
H, W, hidden, L, B = 128, 1024, [64, 128, 256, 512], 256, 2
torch.manual_seed(0)
for dtype in (torch.bfloat16, torch.float32):
    model = tv.VanillaVAE(1, L, (H, W), hidden_dims=list(hidden), compute_dtype=dtype).cuda().train()
    x = (torch.rand(B, 1, H, W) < 0.1).float().cuda()
    eps = torch.randn(B, L).cuda()
    torch.cuda.synchronize(); t0 = time.time()
    out = model(x, eps=eps); losses = model.loss(out); losses["loss"].backward()
    torch.cuda.synchronize(); t1 = time.time()
    print(dtype, "loss", float(losses["loss"]), "time", t1 - t0, "launches", tv.ops.LAUNCHES, "mem GB", torch.cuda.max_memory_allocated() / 1e9)
    del model

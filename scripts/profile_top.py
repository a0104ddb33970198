"""One launch of each dominant kernel of the B=512 bf16 step (for `ncu --set full`)."""
import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_vae_b200 as tv
ops = tv.ops
B = int(os.environ.get("B", "512")); H = W = 128; C = 32
dev = "cuda"
y = torch.randn(B, H, W, C, device=dev).bfloat16()
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
w_last = torch.randn(1, C, 3, 3, device=dev) / math.sqrt(9 * C); b_last = torch.zeros(1, device=dev)
dl = (torch.randn(B, 1, H, W, device=dev) / (B * H * W)).bfloat16()
vec = torch.stack([torch.ones(C), torch.zeros(C), torch.zeros(C), torch.ones(C)]).to(dev)
for rep in range(2):  # first round warms up (attribute setup, tensor maps)
    ops.final_conv_fwd(y, vec, beta, w_last, b_last)
    ops.final_block_backward(dl, y, vec, gamma, beta, w_last, True)
    x64 = torch.randn(B, 64, 64, 32, device=dev).bfloat16()
    w = torch.randn(32, 32, 3, 3, device=dev) / math.sqrt(288)
    cw = ops.ConvWeights(w, torch.bfloat16, 2)
    st = torch.empty(64, dtype=torch.float64, device=dev)
    ops.conv_data_grad(x64, cw, torch.zeros(32, device=dev), st)      # final convT forward (rw mode 1)
    ops.conv_forward(y, cw, None)                                     # its data gradient (rw mode 0)
    ops.conv_weight_grad(y, x64, 2, 32, 32, H, W)                     # its weight gradient
    a64 = torch.randn(B, 64, 64, 32, device=dev).bfloat16()
    ops.bn_apply_lrelu(a64, vec, beta)
    ops.bn_lrelu_backward(a64, x64, vec, gamma, beta, True)
torch.cuda.synchronize()
print("ok")

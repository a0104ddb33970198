import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_vae_b200 as tv
from oracle import vae_oracle as vo
H, W, hidden, L, B = 128, 1024, [64, 128, 256, 512], 256, 2
torch.set_num_threads(16)
params = vo.closed_form_params(vo.param_shapes(1, L, H, W, list(hidden)))
x, eps = vo.closed_form_inputs(B, H, W, L)
r = vo.step(params, x, eps, kld_weight=1.0, dtype=torch.float64, need_grads=False)
model = tv.VanillaVAE(1, L, (H, W), hidden_dims=list(hidden), compute_dtype=torch.float32)
model.load_state_dict({**model.state_dict(), **params}); model = model.cuda().train()
out = model(x.cuda(), eps=eps.cuda()); losses = model.loss(out)
for k in ("loss", "reconstruction_loss", "kld_loss"):
    print(k, float(losses[k]), float(r[k]))
for k, got in (("mu", out["encoded"]["mu"]), ("log_var", out["encoded"]["log_var"]), ("latents", out["latents"]), ("pre_latents", out["encoded"]["pre_latents"]), ("output", out["output"]), ("logits", out["logits"])):
    print(k, vo.rel_err(got, r[k].reshape(got.shape)), float(got.abs().max()))

"""CPU emulation of the CUDA path's bf16 STORAGE points.  TEST INFRASTRUCTURE ONLY.

``vae_oracle.step`` is the fp32/fp64 restatement of the reference
(midi_autoencoder/models.py:185-225, train.py:650).  This file restates the same
step with one change: every tensor the B200 path STORES in bf16 is rounded to
bf16 at that point (values and, on the way back, gradients) while all arithmetic
stays fp32 -- conv outputs y, activations, decoder_input's output, logits, and
the conv weights' operand copies.  It answers a question the parity tests need:
how far from the fp32 reference is ANY faithful bf16-activation implementation
of this network?  (At initialisation the answer is ~10 % on several gradients:
rounding the conv weights alone to bf16 moves them that much -- the train-mode
BatchNorm backward cancels most of the incoming gradient, so the gradient is an
ill-conditioned function of every tensor in the graph.)
"""

from __future__ import annotations

import torch
import torch.nn.functional as F

from . import vae_oracle as vo


class _RoundBoth(torch.autograd.Function):
    """bf16 round-trip of the value (forward) and of its gradient (backward)."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


class _RoundValue(torch.autograd.Function):
    """bf16 operand copy of an fp32 master weight: gradient stays fp32."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


def step(params, x, eps, kld_weight: float = 1.0):
    """Forward + ELBO + gradients with bf16 storage emulated; fp32 arithmetic."""
    R, RW = _RoundBoth.apply, _RoundValue.apply
    P = {k: v.detach().clone().float().requires_grad_(True) for k, v in params.items()}
    Bf = vo.fresh_buffers(params)
    n_enc, n_dec = vo.count_blocks(P)
    h = x.float()
    for i in range(n_enc):
        first = i == 0  # the fp32 piano-roll feeds conv0 unrounded
        y = R(F.conv2d(h, RW(P[f"encoder.{i}.0.weight"]), P[f"encoder.{i}.0.bias"], stride=2, padding=1))
        h = R(vo._bn_lrelu(y, f"encoder.{i}.1", P, Bf, True))
        del first
    feat = h.shape[1:]
    pre = h.flatten(1)
    mu = F.linear(pre, P["fc_mu.weight"], P["fc_mu.bias"])
    log_var = F.linear(pre, P["fc_var.weight"], P["fc_var.bias"])
    z = vo.reparameterize(mu, log_var, eps.float())
    h = R(F.linear(z, P["decoder_input.weight"], P["decoder_input.bias"])).view(-1, *feat)
    blocks = [f"decoder.{i}" for i in range(n_dec)] + ["final_layer"]
    for name in blocks:
        y = R(F.conv_transpose2d(h, RW(P[f"{name}.0.weight"]), P[f"{name}.0.bias"], stride=2, padding=1, output_padding=1))
        h = R(vo._bn_lrelu(y, f"{name}.1", P, Bf, True))
    logits = R(F.conv2d(h, RW(P["final_layer.3.weight"]), P["final_layer.3.bias"], stride=1, padding=1))
    output = torch.sigmoid(logits)
    loss, recon, kld = vo.elbo(output, x.float(), mu, log_var, kld_weight)
    loss.backward()
    return {
        "output": output.detach(), "mu": mu.detach(), "log_var": log_var.detach(), "latents": z.detach(),
        "loss": loss.detach(), "reconstruction_loss": recon, "kld_loss": kld,
        "grads": {k: v.grad.detach() for k, v in P.items()},
    }

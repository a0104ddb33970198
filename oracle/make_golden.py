"""Generate tests/golden/*.npz by running the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

Runs only in the build container, where ``/root/reference`` exists:

    python oracle/make_golden.py

It imports ``/root/reference/midi_autoencoder/models.py`` (flat import of
``types_helpers``, models.py:4), loads deterministic closed-form weights
(``vae_oracle.closed_form_params``) into it, runs forward (models.py:185-188),
``loss`` (:190-225) and ``.backward()`` (train.py:650) and stores what came out.

Cases
-----
asis_32      the reference class exactly as shipped: 1x32x32, default widths,
             L=10 -- the only shape it accepts (models.py:33,166).
shim_32      the same through the generalising shim; asserted BIT-IDENTICAL to
             asis_32 here, which proves the shim changes no arithmetic.
shim_128     shim, 1x128x128 default widths (BASELINE configs 1-3 shape).
shim_rect    shim, 1x32x64 (non-square), widths [8,16,32,64], L=6,
             kld_weight 0.00025 (configs/vae.yaml:19).
eval_32      eval()-mode forward after one training step (running-stat BN).

To stay small each fixture stores inputs in full (bit-packed x, eps), small
outputs in full, and for every large tensor (parameter gradients, images) its
sum, L2 norm and a fixed strided sample.  Weights are closed-form, so they are
regenerated, not stored.
"""

from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
REF = "/root/reference/midi_autoencoder"

from oracle import vae_oracle as vo  # noqa: E402


def import_reference():
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not present; fixtures can only be made in the build container")
    sys.path.insert(0, REF)
    import models  # type: ignore  # noqa

    return models.VanillaVAE


def build_reference(VanillaVAE, height, width, hidden_dims, latent, kld_weight, shim: bool):
    """Instantiate the reference; with ``shim`` generalise its two literals
    (SURVEY.md §8c rule) without touching any other arithmetic."""
    with contextlib.redirect_stdout(io.StringIO()):  # ctor prints init stats (models.py:238-247)
        model = VanillaVAE(1, latent, input_dim=height, hidden_dims=list(hidden_dims), kld_weight=kld_weight)
    if not shim:
        return model
    n = len(hidden_dims)
    h, w = height, width
    for _ in range(n):
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    c_last = hidden_dims[-1]
    flat = c_last * h * w
    model.flattened_size = flat
    model.fc_mu = torch.nn.Linear(flat, latent)
    model.fc_var = torch.nn.Linear(flat, latent)
    model.decoder_input = torch.nn.Linear(latent, flat)

    def decode(z, _m=model, _shape=(c_last, h, w)):
        t = _m.decoder_input(z).view(-1, *_shape)
        return _m.final_layer(_m.decoder(t))

    model.decode = decode
    return model


def run_reference(model, params, x, eps, training=True):
    model.load_state_dict({**model.state_dict(), **params})
    model.reparameterize = lambda mu, lv: eps * torch.exp(0.5 * lv) + mu
    model.train(training)
    for p in model.parameters():
        p.grad = None
    out = model.forward(x)
    losses = model.loss(out)
    grads = {}
    if training:
        losses["loss"].backward()
        grads = {k: v.grad.detach().clone() for k, v in model.named_parameters()}
    bufs = {k: v.detach().clone() for k, v in model.named_buffers()}
    return out, losses, grads, bufs


SAMPLE = 257  # stride of the element sample kept for large tensors


def summarise(t: torch.Tensor):
    f = t.detach().double().flatten()
    return np.array([float(f.sum()), float(f.norm())]), f[::SAMPLE][:64].numpy().astype(np.float32)


def pack_case(name, height, width, hidden, latent, kldw, batch, out, losses, grads, bufs, x, eps):
    d = {
        "meta": np.array([height, width, latent, batch] + list(hidden), dtype=np.int64),
        "kld_weight": np.array(kldw, dtype=np.float64),
        "x_bits": np.packbits(x.numpy().astype(np.uint8).reshape(-1)),
        "eps": eps.numpy(),
        "loss": losses["loss"].detach().numpy(),
        "reconstruction_loss": losses["reconstruction_loss"].numpy(),
        "kld_loss": losses["kld_loss"].numpy(),
        "mu": out["encoded"]["mu"].detach().numpy(),
        "log_var": out["encoded"]["log_var"].detach().numpy(),
        "latents": out["latents"].detach().numpy(),
    }
    for key, t in (("output", out["output"]), ("pre_latents", out["encoded"]["pre_latents"])):
        s, e = summarise(t)
        d[f"{key}.stat"], d[f"{key}.samp"] = s, e
    for k, g in grads.items():
        s, e = summarise(g)
        d[f"grad/{k}.stat"], d[f"grad/{k}.samp"] = s, e
    for k, b in bufs.items():
        d[f"buf/{k}"] = b.numpy()
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", f"{name}.npz")
    np.savez_compressed(path, **d)
    print(f"wrote {path}: {os.path.getsize(path)} bytes, loss={float(losses['loss']):.8f}")


def main():
    torch.set_num_threads(1)  # summation order fixed -> fixtures reproducible
    VanillaVAE = import_reference()
    cases = [
        # name, H, W, hidden, L, kld_weight, batch, shim
        ("asis_32", 32, 32, [32, 64, 128, 256], 10, 1.0, 8, False),
        ("shim_32", 32, 32, [32, 64, 128, 256], 10, 1.0, 8, True),
        ("shim_128", 128, 128, [32, 64, 128, 256], 10, 1.0, 2, True),
        ("shim_rect", 32, 64, [8, 16, 32, 64], 6, 0.00025, 4, True),
    ]
    results = {}
    for name, H, W, hidden, L, kldw, B, shim in cases:
        shapes = vo.param_shapes(1, L, H, W, list(hidden))
        params = vo.closed_form_params(shapes)
        x, eps = vo.closed_form_inputs(B, H, W, L)
        model = build_reference(VanillaVAE, H, W, hidden, L, kldw, shim)
        assert {k: tuple(v.shape) for k, v in model.named_parameters()} == shapes, name
        out, losses, grads, bufs = run_reference(model, params, x, eps)
        results[name] = (out, losses, grads, bufs)
        if name != "shim_32":
            pack_case(name, H, W, hidden, L, kldw, B, out, losses, grads, bufs, x, eps)
        if name == "asis_32":
            # eval-mode forward on the SAME module (running stats now updated once)
            model.eval()
            with torch.no_grad():
                out_e = model.forward(x)
                loss_e = model.loss(out_e)
            pack_case("eval_32", H, W, hidden, L, kldw, B, out_e, loss_e, {}, bufs, x, eps)

    # the shim must not change a single bit at the one shape the reference accepts
    a, s = results["asis_32"], results["shim_32"]
    assert torch.equal(a[0]["output"], s[0]["output"])
    assert torch.equal(a[1]["loss"], s[1]["loss"])
    for k in a[2]:
        assert torch.equal(a[2][k], s[2][k]), k
    for k in a[3]:
        assert torch.equal(a[3][k], s[3][k]), k
    print("shim == as-is reference at 1x32x32: bit-identical (outputs, loss, 40 grads, 24 buffers)")


if __name__ == "__main__":
    main()

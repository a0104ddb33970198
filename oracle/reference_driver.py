"""Drive the UNMODIFIED reference ``VanillaVAE`` class.  TEST / BASELINE INFRASTRUCTURE ONLY.

Used by ``bench.py``'s reference arms (``--impl reference``, ``cpu_baseline``,
``gpu_eager_baseline``) and by the drop-in interoperability tests; nothing under
``torch_vae_b200`` may import it.

Where the reference comes from: ``baseline/_ref/midi_autoencoder`` (installed once with
``pip install --no-index --no-deps --target baseline/_ref`` from a copy of
``/root/reference``; git-ignored, travels to the GPU box) or, in the build container,
``/root/reference/midi_autoencoder``.  The reference is a flat-import script directory
(``from types_helpers import ...``, models.py:4), so its directory is put on ``sys.path``
only while ``models`` / ``evaluation`` / ``utils`` are imported.

The reference class runs as shipped at 32x32 only (models.py:33,166).  For any other input
extent the two literals are generalised by the rule of SURVEY.md section 8c -- construct the
reference class, swap the three ``nn.Linear`` for correctly sized ones and override ``decode``
with the same three calls -- which ``oracle/make_golden.py`` proves bit-identical to the
unmodified class at 32x32.
"""

from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
from typing import List, Optional

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = [os.path.join(ROOT, "baseline", "_ref", "midi_autoencoder"), "/root/reference/midi_autoencoder"]


def reference_dir() -> Optional[str]:
    for d in CANDIDATES:
        if os.path.isfile(os.path.join(d, "models.py")):
            return d
    return None


def available() -> bool:
    return reference_dir() is not None


def import_module(name: str):
    """Import one of the reference's flat modules (``models``, ``evaluation``, ``utils`` ...)."""
    d = reference_dir()
    if d is None:
        raise RuntimeError("reference not installed: baseline/_ref/midi_autoencoder and /root/reference are both absent")
    cached = sys.modules.get(name)
    if cached is not None and os.path.dirname(getattr(cached, "__file__", "") or "") == d:
        return cached
    sys.path.insert(0, d)
    try:
        if name in sys.modules:  # a same-named foreign module (e.g. HF ``datasets``) must not be returned
            del sys.modules[name]
        with contextlib.redirect_stdout(io.StringIO()):
            return importlib.import_module(name)
    finally:
        sys.path.remove(d)


def build_model(height: int, width: int, hidden_dims: List[int], latent: int, kld_weight: float = 1.0):
    """The reference class; generalised (SURVEY.md section 8c rule) unless the input is 32x32."""
    VanillaVAE = import_module("models").VanillaVAE
    with contextlib.redirect_stdout(io.StringIO()):  # the ctor prints init statistics (models.py:238-247)
        model = VanillaVAE(1, latent, input_dim=height, hidden_dims=list(hidden_dims), kld_weight=kld_weight)
    if (height, width) == (32, 32) and len(hidden_dims) == 4:
        return model
    h, w = height, width
    for _ in hidden_dims:
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    c_last = model.hidden_dims[0]  # the ctor reversed the list in place (models.py:60)
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


class ReferenceTrainStep:
    """train.py:635,642,644,650,656 in that order with the reference's optimizer construction
    (AdamW over ``model.encoder`` / ``model.decoder`` only, train.py:210-228; lr = 0.01 * batch / 128,
    train.py:201; weight decay 0, the CLI default).  ``autocast`` = torch.bfloat16 runs forward and
    loss under ``torch.autocast`` the way a user would ask eager PyTorch for bf16."""

    def __init__(self, height, width, hidden_dims, latent, batch, device="cpu", autocast=None, world: int = 1):
        torch.manual_seed(0)
        self.model = build_model(height, width, hidden_dims, latent).to(device).train()
        lr = 0.01 * batch * world / 128
        groups = [{"params": self.model.encoder.parameters(), "lr": lr, "name": "encoder"},
                  {"params": self.model.decoder.parameters(), "lr": lr, "name": "decoder"}]
        self.opt = torch.optim.AdamW(groups, lr=lr, weight_decay=0.0)
        self.autocast = autocast
        self.device = torch.device(device)

    def __call__(self, x):
        m = self.model
        if self.autocast is not None:
            with torch.autocast(self.device.type, dtype=self.autocast):
                out = m.forward(x)
            # F.binary_cross_entropy is not autocast-safe: the loss runs on fp32 probabilities, as eager users do
            out["output"] = out["output"].float()
            out["encoded"]["mu"] = out["encoded"]["mu"].float()
            out["encoded"]["log_var"] = out["encoded"]["log_var"].float()
        else:
            out = m.forward(x)
        self.opt.zero_grad()
        losses = m.loss(out)
        losses["loss"].backward()
        self.opt.step()
        return losses["loss"].detach()

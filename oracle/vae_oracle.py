"""CPU oracle for the VanillaVAE train-step hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch, CPU, functional restatement of the arithmetic the
reference performs in ``midi_autoencoder/models.py`` (forward :185-188, encode
:107-145, reparameterize :177-183, decode :147-175, loss :190-225) driven the
way ``midi_autoencoder/train.py:635-650`` drives it.  Nothing under
``torch_vae_b200`` may import it: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs do, and there only
as the checker or as the timed CPU baseline, never as the product path.

Parity status: the reference ships no tests, golden vectors or fixtures for
this path.  The oracle is therefore pinned against OUTPUTS
OF THE REFERENCE ITSELF: ``oracle/make_golden.py`` imports the unmodified
``/root/reference/midi_autoencoder/models.py`` in the build container, runs it
on deterministic closed-form weights / inputs / eps, and commits the results
under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file
against those fixtures on every CPU run.

The arithmetic itself lives in a third-party dependency of the reference,
PyTorch ATen (``torch==2.5.1`` pinned in the reference's requirements.txt:24;
this image has 2.11.0).  The restatement calls the same ATen ops
(``conv2d``, ``conv_transpose2d``, ``batch_norm``, ``leaky_relu``, ``linear``,
``sigmoid``, ``binary_cross_entropy``) in the same order, functionally, on a
dict of tensors keyed by the reference's ``state_dict`` names, and generalises
the two literals that pin the reference to 32x32 inputs (models.py:33 and
:166) by the rule in SURVEY.md §8c: h, w = H/16, W/16.
"""

from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

BN_EPS = 1e-5  # nn.BatchNorm2d default (models.py:46,69,78)
BN_MOMENTUM = 0.1  # nn.BatchNorm2d default
LRELU_SLOPE = 0.01  # nn.LeakyReLU default (models.py:47,70,79)


# --------------------------------------------------------------------------
# structure helpers
# --------------------------------------------------------------------------
def count_blocks(params: Dict[str, Tensor]) -> Tuple[int, int]:
    """Number of encoder / decoder blocks encoded in the state_dict keys."""
    n_enc = len({k.split(".")[1] for k in params if k.startswith("encoder.")})
    n_dec = len({k.split(".")[1] for k in params if k.startswith("decoder.")})
    return n_enc, n_dec


def param_shapes(
    in_channels: int, latent_dim: int, height: int, width: int, hidden_dims: List[int]
) -> Dict[str, Tuple[int, ...]]:
    """Shapes of the 40 parameters (for 4 hidden dims) in reference state_dict
    order (models.py:40-83)."""
    n = len(hidden_dims)
    h, w = height, width
    for _ in range(n):
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    flat = hidden_dims[-1] * h * w
    shapes: Dict[str, Tuple[int, ...]] = {}
    cin = in_channels
    for i, c in enumerate(hidden_dims):
        shapes[f"encoder.{i}.0.weight"] = (c, cin, 3, 3)
        shapes[f"encoder.{i}.0.bias"] = (c,)
        shapes[f"encoder.{i}.1.weight"] = (c,)
        shapes[f"encoder.{i}.1.bias"] = (c,)
        cin = c
    shapes["fc_mu.weight"] = (latent_dim, flat)
    shapes["fc_mu.bias"] = (latent_dim,)
    shapes["fc_var.weight"] = (latent_dim, flat)
    shapes["fc_var.bias"] = (latent_dim,)
    shapes["decoder_input.weight"] = (flat, latent_dim)
    shapes["decoder_input.bias"] = (flat,)
    rev = list(reversed(hidden_dims))
    for i in range(n - 1):
        shapes[f"decoder.{i}.0.weight"] = (rev[i], rev[i + 1], 3, 3)
        shapes[f"decoder.{i}.0.bias"] = (rev[i + 1],)
        shapes[f"decoder.{i}.1.weight"] = (rev[i + 1],)
        shapes[f"decoder.{i}.1.bias"] = (rev[i + 1],)
    c = rev[-1]
    shapes["final_layer.0.weight"] = (c, c, 3, 3)
    shapes["final_layer.0.bias"] = (c,)
    shapes["final_layer.1.weight"] = (c,)
    shapes["final_layer.1.bias"] = (c,)
    shapes["final_layer.3.weight"] = (1, c, 3, 3)
    shapes["final_layer.3.bias"] = (1,)
    return shapes


def bn_names(params: Dict[str, Tensor]) -> List[str]:
    """Prefixes of the BatchNorm2d modules, in execution order."""
    n_enc, n_dec = count_blocks(params)
    return (
        [f"encoder.{i}.1" for i in range(n_enc)]
        + [f"decoder.{i}.1" for i in range(n_dec)]
        + ["final_layer.1"]
    )


def fresh_buffers(params: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """BatchNorm buffers as nn.BatchNorm2d creates them (mean 0, var 1, count 0)."""
    bufs: Dict[str, Tensor] = {}
    for p in bn_names(params):
        c = params[p + ".weight"].numel()
        dt = params[p + ".weight"].dtype
        bufs[p + ".running_mean"] = torch.zeros(c, dtype=dt)
        bufs[p + ".running_var"] = torch.ones(c, dtype=dt)
        bufs[p + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return bufs


# --------------------------------------------------------------------------
# deterministic closed-form tensors (no RNG: identical on every machine)
# --------------------------------------------------------------------------
def _fill(shape, key: float, scale: float, offset: float = 0.0) -> Tensor:
    n = int(math.prod(shape)) if len(shape) else 1
    idx = torch.arange(n, dtype=torch.float64)
    u = torch.remainder(idx * 0.6180339887498949 + key * 0.7548776662466927 + 0.25, 1.0)
    return ((u - 0.5) * 2.0 * scale + offset).reshape(shape).to(torch.float32)


def closed_form_params(shapes: Dict[str, Tuple[int, ...]]) -> Dict[str, Tensor]:
    """Deterministic weights with xavier-like magnitudes; BN gamma near 1, beta
    and all biases small but NON-zero so every bias path is exercised."""
    out: Dict[str, Tensor] = {}
    for j, (name, shp) in enumerate(shapes.items()):
        key = float(j + 1)
        if name.endswith(".weight") and len(shp) == 4:
            fan = shp[1] * 9 + shp[0] * 9
            out[name] = _fill(shp, key, math.sqrt(6.0 / fan))
        elif name.endswith(".weight") and len(shp) == 2:
            out[name] = _fill(shp, key, math.sqrt(6.0 / (shp[0] + shp[1])))
        elif name.endswith(".weight"):  # BN gamma
            out[name] = _fill(shp, key, 0.25, 1.0)
        else:
            out[name] = _fill(shp, key, 0.05)
    return out


def closed_form_inputs(batch: int, height: int, width: int, latent: int, key: float = 99.0):
    """Binary piano-roll (about 10 % density) and eps, closed form."""
    u = _fill((batch, 1, height, width), key, 0.5, 0.5)
    x = (u < 0.10).to(torch.float32)
    # Box-Muller on two low-discrepancy streams -> N(0,1)-like eps
    u1 = _fill((batch, latent), key + 1.0, 0.4999, 0.5).double()
    u2 = _fill((batch, latent), key + 2.0, 0.5, 0.5).double()
    eps = (torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2.0 * math.pi * u2)).to(torch.float32)
    return x, eps


# --------------------------------------------------------------------------
# forward pieces (functional)
# --------------------------------------------------------------------------
def _bn_lrelu(y: Tensor, prefix: str, P, Bf, training: bool) -> Tensor:
    """BatchNorm2d + LeakyReLU as run by the reference (models.py:46-47).

    In training mode F.batch_norm normalises with the biased batch variance
    and updates the running buffers in place with the unbiased one.
    """
    rm, rv = Bf[prefix + ".running_mean"], Bf[prefix + ".running_var"]
    if training:
        Bf[prefix + ".num_batches_tracked"] += 1
    u = F.batch_norm(
        y, rm, rv, P[prefix + ".weight"], P[prefix + ".bias"], training, BN_MOMENTUM, BN_EPS
    )
    return F.leaky_relu(u, LRELU_SLOPE)


def encode(P, Bf, x: Tensor, training: bool = True):
    """models.py:107-145: conv stack -> flatten (C-major) -> fc_mu / fc_var."""
    n_enc, _ = count_blocks(P)
    h = x
    for i in range(n_enc):
        h = F.conv2d(h, P[f"encoder.{i}.0.weight"], P[f"encoder.{i}.0.bias"], stride=2, padding=1)
        h = _bn_lrelu(h, f"encoder.{i}.1", P, Bf, training)
    feat_shape = h.shape[1:]
    pre = h.flatten(start_dim=1)
    mu = F.linear(pre, P["fc_mu.weight"], P["fc_mu.bias"])
    log_var = F.linear(pre, P["fc_var.weight"], P["fc_var.bias"])
    return mu, log_var, pre, feat_shape


def reparameterize(mu: Tensor, log_var: Tensor, eps: Tensor) -> Tensor:
    """models.py:181-183 with eps injected instead of randn_like."""
    std = torch.exp(0.5 * log_var)
    return eps * std + mu


def decode_logits(P, Bf, z: Tensor, feat_shape, training: bool = True) -> Tensor:
    """models.py:147-175 up to (not including) the Sigmoid; the view literal
    (:166) is generalised to the encoder's own feature shape."""
    _, n_dec = count_blocks(P)
    h = F.linear(z, P["decoder_input.weight"], P["decoder_input.bias"])
    h = h.view(-1, *feat_shape)
    for i in range(n_dec):
        h = F.conv_transpose2d(
            h, P[f"decoder.{i}.0.weight"], P[f"decoder.{i}.0.bias"],
            stride=2, padding=1, output_padding=1,
        )
        h = _bn_lrelu(h, f"decoder.{i}.1", P, Bf, training)
    h = F.conv_transpose2d(
        h, P["final_layer.0.weight"], P["final_layer.0.bias"], stride=2, padding=1, output_padding=1
    )
    h = _bn_lrelu(h, "final_layer.1", P, Bf, training)
    return F.conv2d(h, P["final_layer.3.weight"], P["final_layer.3.bias"], stride=1, padding=1)


def elbo(output: Tensor, x: Tensor, mu: Tensor, log_var: Tensor, kld_weight: float):
    """models.py:208-225: mean BCE on probabilities + kld_weight * KL; the
    returned kld_loss is sign-flipped exactly as the reference does (:224)."""
    recon = F.binary_cross_entropy(output, x)
    kld = -0.5 * torch.mean(torch.sum(1 + log_var - mu**2 - torch.exp(log_var), dim=-1))
    loss = recon + kld_weight * kld
    return loss, recon.detach(), -kld.detach()


# --------------------------------------------------------------------------
# one whole step: forward + loss + backward (train.py:635,644,650)
# --------------------------------------------------------------------------
def feature_shape(P, height: int, width: int) -> Tuple[int, int, int]:
    n_enc, _ = count_blocks(P)
    h, w = height, width
    for _ in range(n_enc):
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    return (P[f"encoder.{n_enc - 1}.0.weight"].shape[0], h, w)


def step(
    params: Dict[str, Tensor],
    x: Tensor,
    eps: Tensor,
    kld_weight: float = 1.0,
    buffers: Optional[Dict[str, Tensor]] = None,
    training: bool = True,
    dtype: torch.dtype = torch.float32,
    need_grads: bool = True,
):
    """Forward, ELBO and gradients of every parameter.

    Returns a dict with ``output`` (probabilities), ``logits``, ``mu``,
    ``log_var``, ``latents``, ``pre_latents``, the three loss scalars of
    ``LossOutput`` (types_helpers.py:32-37), ``grads`` (name -> tensor) and
    ``buffers`` (updated BatchNorm buffers).
    """
    P = {k: v.detach().to(dtype).clone().requires_grad_(need_grads) for k, v in params.items()}
    if buffers is None:
        buffers = fresh_buffers(params)
    Bf = {
        k: (v.clone() if v.dtype == torch.long else v.detach().to(dtype).clone())
        for k, v in buffers.items()
    }
    xx = x.to(dtype)
    mu, log_var, pre, feat_shape = encode(P, Bf, xx, training)
    z = reparameterize(mu, log_var, eps.to(dtype))
    logits = decode_logits(P, Bf, z, feat_shape, training)
    output = torch.sigmoid(logits)  # models.py:81
    loss, recon, kld_flipped = elbo(output, xx, mu, log_var, kld_weight)
    grads: Dict[str, Tensor] = {}
    if need_grads:
        loss.backward()
        grads = {k: v.grad.detach() for k, v in P.items()}
    return {
        "output": output.detach(),
        "logits": logits.detach(),
        "mu": mu.detach(),
        "log_var": log_var.detach(),
        "latents": z.detach(),
        "pre_latents": pre.detach(),
        "loss": loss.detach(),
        "reconstruction_loss": recon,
        "kld_loss": kld_flipped,
        "grads": grads,
        "buffers": Bf,
    }


def decode_only(params, z: Tensor, height: int, width: int, buffers=None, training: bool = False):
    """models.py:250-263 (``sample``) minus the host randn: z -> probabilities."""
    P = {k: v.detach() for k, v in params.items()}
    Bf = {k: v.clone() for k, v in (buffers or fresh_buffers(params)).items()}
    with torch.no_grad():
        logits = decode_logits(P, Bf, z, feature_shape(P, height, width), training)
        return torch.sigmoid(logits)


# --------------------------------------------------------------------------
# error metric used by every parity test (SURVEY.md §8c "Tolerance mechanics")
# --------------------------------------------------------------------------
def rel_err(a: Tensor, b: Tensor, floor: float = 0.0) -> float:
    """Norm-wise relative error ||a-b|| / max(||b||, floor)."""
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    den = max(float(b.norm()), floor)
    if den == 0.0:
        return float((a - b).norm())
    return float((a - b).norm()) / den


def grad_floor(grads: Dict[str, Tensor], name: str) -> float:
    """Absolute floor for the norm-wise error of ``grads[name]``.

    The eight conv biases that feed a BatchNorm have an analytically zero
    gradient (reference fp32 value: 1e-7..1e-9 rounding noise);
    relative error is undefined there, so their error is measured against the
    gradient norm of the weight of the same module instead.
    """
    if name.endswith(".bias"):
        w = name[: -len("bias")] + "weight"
        if w in grads:
            return float(grads[w].detach().double().norm())
    return 0.0


# --------------------------------------------------------------------------
# synthetic workload of BASELINE.md §3 (seeded torch.Generator)
# --------------------------------------------------------------------------
def synthetic_batch(batch: int, height: int, width: int, latent: int, seed: int = 1234):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(batch, 1, height, width, generator=g) < 0.10).float()
    eps = torch.randn(batch, latent, generator=g)
    return x, eps


def xavier_params(
    in_channels: int, latent: int, height: int, width: int, hidden_dims: List[int], seed: int = 0
) -> Dict[str, Tensor]:
    """Random-init weights with the same distributions the reference's ctor
    yields (models.py:227-236 xavier for Conv2d inside the Sequentials, torch
    defaults elsewhere) -- used for benchmarking only, where the exact draw
    does not matter.  Parity tests use ``closed_form_params``."""
    g = torch.Generator().manual_seed(seed)
    shapes = param_shapes(in_channels, latent, height, width, list(hidden_dims))
    out: Dict[str, Tensor] = {}
    for name, shp in shapes.items():
        if name.endswith(".weight") and len(shp) == 4:
            rf = 9
            bound = math.sqrt(6.0 / (shp[1] * rf + shp[0] * rf))
            out[name] = (torch.rand(shp, generator=g) * 2 - 1) * bound
        elif name.endswith(".weight") and len(shp) == 2:
            bound = 1.0 / math.sqrt(shp[1])
            out[name] = (torch.rand(shp, generator=g) * 2 - 1) * bound
        elif name.endswith(".weight"):
            out[name] = torch.ones(shp)
        elif name.startswith(("fc_", "decoder_input")):
            bound = 1.0 / math.sqrt(shapes[name.replace(".bias", ".weight")][1])
            out[name] = (torch.rand(shp, generator=g) * 2 - 1) * bound
        else:
            out[name] = torch.zeros(shp)
    return out


# --------------------------------------------------------------------------
# the reference's CPU training step, for the timed CPU baseline of bench.py
# --------------------------------------------------------------------------
class CpuTrainStep:
    """train.py:635,642,644,650,656 in that order on the host CPU: forward,
    zero_grad, loss, backward, optimizer.step -- AdamW over the encoder and
    decoder parameter groups ONLY (train.py:210-228, "quirk A" in SURVEY.md
    §3.1), lr = 0.01 * batch / 128 (train.py:201), weight decay 0 as in the
    reference CLI default.  Functional ops from this file, persistent leaf
    parameters, fresh eps from torch.randn like models.py:182."""

    def __init__(self, params: Dict[str, Tensor], batch: int, kld_weight: float = 1.0, lr_relative: float = 0.01):
        self.P = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
        self.Bf = fresh_buffers(params)
        self.kld_weight = kld_weight
        lr = lr_relative * batch / 128
        enc = [v for k, v in self.P.items() if k.startswith("encoder.")]
        dec = [v for k, v in self.P.items() if k.startswith("decoder.")]
        self.opt = torch.optim.AdamW(
            [{"params": enc, "lr": lr, "name": "encoder"}, {"params": dec, "lr": lr, "name": "decoder"}],
            lr=lr, weight_decay=0.0,
        )

    def __call__(self, x: Tensor, eps: Optional[Tensor] = None):
        mu, log_var, pre, feat_shape = encode(self.P, self.Bf, x, True)
        z = reparameterize(mu, log_var, torch.randn_like(mu) if eps is None else eps)
        output = torch.sigmoid(decode_logits(self.P, self.Bf, z, feat_shape, True))
        for p in self.P.values():
            p.grad = None
        loss, recon, kld = elbo(output, x, mu, log_var, self.kld_weight)
        loss.backward()
        self.opt.step()
        return loss.detach()

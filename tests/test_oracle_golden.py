"""The CPU oracle must reproduce what the UNMODIFIED reference produced.

Fixtures under tests/golden/ were written by oracle/make_golden.py from
/root/reference/midi_autoencoder/models.py (forward :185-188, loss :190-225,
backward train.py:650).  This pins the oracle; the GPU parity tests then pin
the CUDA path to the oracle.
"""
import os

import numpy as np
import pytest
import torch

from oracle import vae_oracle as vo

SAMPLE = 257
CASES = ["asis_32", "shim_128", "shim_rect"]


def load_case(golden_dir, name):
    d = np.load(os.path.join(golden_dir, f"{name}.npz"))
    meta = d["meta"].tolist()
    H, W, L, B = meta[:4]
    hidden = meta[4:]
    x = torch.from_numpy(np.unpackbits(d["x_bits"])[: B * H * W].reshape(B, 1, H, W).astype(np.float32))
    eps = torch.from_numpy(d["eps"])
    params = vo.closed_form_params(vo.param_shapes(1, L, H, W, hidden))
    return d, params, x, eps, float(d["kld_weight"])


def check_big(d, key, t, tol):
    f = t.double().flatten()
    s, n = d[f"{key}.stat"]
    assert abs(float(f.norm()) - n) <= tol * max(n, 1e-6), key
    samp = f[::SAMPLE][:64].float().numpy()
    ref = d[f"{key}.samp"]
    assert np.linalg.norm(samp - ref) <= tol * max(np.linalg.norm(ref), 1e-6) + 1e-9, key
    assert abs(float(f.sum()) - s) <= tol * max(n, 1e-6) * 10 + 1e-7, key


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_train_step(golden_dir, name):
    torch.set_num_threads(1)
    d, params, x, eps, kldw = load_case(golden_dir, name)
    # the closed-form inputs are regenerable too: check the generator is stable
    H, W, L, B = d["meta"].tolist()[:4]
    x2, eps2 = vo.closed_form_inputs(B, H, W, L)
    assert torch.equal(x, x2) and torch.equal(eps, eps2)

    r = vo.step(params, x, eps, kld_weight=kldw)
    tol = 2e-6
    for k in ("loss", "reconstruction_loss", "kld_loss"):
        assert abs(float(r[k]) - float(d[k])) <= tol * abs(float(d[k])) + 1e-7, k
    for k in ("mu", "log_var", "latents"):
        assert vo.rel_err(r[k], torch.from_numpy(d[k])) <= tol, k
    check_big(d, "output", r["output"], tol)
    check_big(d, "pre_latents", r["pre_latents"], tol)
    assert len(r["grads"]) == 40
    for k, g in r["grads"].items():
        check_big(d, f"grad/{k}", g, 2e-5)
    for k, b in r["buffers"].items():
        ref = torch.from_numpy(d[f"buf/{k}"])
        if b.dtype == torch.long:
            assert int(b) == int(ref)
        else:
            assert vo.rel_err(b, ref) <= tol, k


def test_oracle_eval_mode(golden_dir):
    torch.set_num_threads(1)
    d, params, x, eps, kldw = load_case(golden_dir, "eval_32")
    bufs = {k[4:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("buf/")}
    r = vo.step(params, x, eps, kld_weight=kldw, buffers=bufs, training=False, need_grads=False)
    assert abs(float(r["loss"]) - float(d["loss"])) <= 2e-6 * abs(float(d["loss"]))
    assert vo.rel_err(r["mu"], torch.from_numpy(d["mu"])) <= 2e-6
    check_big(d, "output", r["output"], 2e-6)


def test_fp64_oracle_close_to_fp32():
    """fp64 restatement separates kernel error from fp32 summation noise."""
    shapes = vo.param_shapes(1, 6, 32, 32, [8, 16, 16, 32])
    params = vo.closed_form_params(shapes)
    x, eps = vo.closed_form_inputs(4, 32, 32, 6)
    r32 = vo.step(params, x, eps)
    r64 = vo.step(params, x, eps, dtype=torch.float64)
    assert vo.rel_err(r32["loss"], r64["loss"]) < 1e-6
    for k in r32["grads"]:
        # conv biases that feed a BatchNorm have an exactly-zero gradient; the
        # fp32 value is rounding noise, so the floor is tied to the magnitude
        # accumulated in the same module's weight gradient.
        floor = vo.grad_floor(r64["grads"], k)
        assert vo.rel_err(r32["grads"][k], r64["grads"][k], floor=floor) < 1e-4, k

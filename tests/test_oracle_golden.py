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


def test_philox_restatement_against_random123_known_answers():
    """oracle/roll_transform_oracle.philox4x32_10 against the Random123 kat_vectors for philox4x32-10."""
    from oracle import roll_transform_oracle as ro

    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, want in kat:
        assert ro.philox4x32_10(ctr, key) == want


def test_roll_transform_oracle_semantics():
    import numpy as np

    from oracle import roll_transform_oracle as ro

    rng = np.random.default_rng(0)
    rolls = (rng.random((3, 128, 160)) < 0.1).astype(np.uint8) * 255
    out, prm = ro.roll_transform(rolls, 128, 128, False, 0, -1.0, 0.5, 1.0, 0, 0)
    assert prm == [(0, 16, 0)] * 3
    assert np.array_equal(out[:, 0], rolls[:, :, 16:144].astype(np.float32) / np.float32(255) - np.float32(0.5))
    out, prm = ro.roll_transform(rolls, 128, 128, True, 5, 0.5, 0.0, 1.0, 9, 4)
    for b, (top, left, shift) in enumerate(prm):
        assert top == 0 and 0 <= left <= 32 and -5 <= shift <= 5
        win = (rolls[b, :, left:left + 128] > 127).astype(np.float32)
        for h in range(128):
            src = h + shift
            want = win[src] if 0 <= src < 128 else np.zeros(128, np.float32)
            assert np.array_equal(out[b, 0, h], want)

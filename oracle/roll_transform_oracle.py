"""CPU restatement (numpy) of the device-side piano-roll transform.  TEST INFRASTRUCTURE ONLY.

What it restates
----------------
* the reference's torchvision stacks ``data_transformations.py:23-46`` ("noaug") and ``:48-72`` ("midi") for
  uint8 rolls whose short side already equals the target size (``Resize`` is then the identity):
  RandomCrop / CenterCrop -> ToImage -> ToDtype(float32, scale=True) (x / 255) -> Normalize(mean, std)
  -> Grayscale (identity for one channel);
* the two augmentations north_star adds and the reference does not have, whose semantics are DEFINED here and
  pinned bit-exactly by tests/test_transforms_gpu.py:
    binarise         v = 1.0 if x / 255 > threshold else 0.0      (before Normalize)
    pitch transpose  out[h] = in[h + shift], zero fill outside the cropped window, shift uniform in
                     [-max_pitch_shift, +max_pitch_shift]
* where the random numbers come from: ONE Philox4x32-10 draw per sample, counter = offset + sample index,
  stream id 0x5eed, key = seed:  x -> crop top, y -> crop left, z -> pitch shift  (each ``% range``).
  CenterCrop follows torchvision: top = int(round((Hs - H) / 2.0)) (round-half-even).

Parity status: Philox4x32-10 is pinned by the Random123 known-answer vectors (tests/test_oracle_golden.py);
the crop / scale / normalise stack is pinned against the reference's own torchvision pipeline run from
baseline/_ref (tests/test_transforms_gpu.py); binarise / transpose have no reference counterpart.
"""
from __future__ import annotations

import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(counter, key):
    """counter: 4 uint32, key: 2 uint32 -> 4 uint32 (Salmon et al., SC'11)."""
    c0, c1, c2, c3 = (int(v) & MASK for v in counter)
    k0, k1 = (int(v) & MASK for v in key)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> 32, p0 & MASK, p1 >> 32, p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def sample_draw(seed: int, counter: int, stream_id: int = 0x5EED):
    return philox4x32_10((counter & MASK, (counter >> 32) & MASK, stream_id, 0), (seed & MASK, (seed >> 32) & MASK))


def center_offset(src: int, dst: int) -> int:
    return int(round((src - dst) / 2.0))  # torchvision.transforms.v2.functional.center_crop


def roll_transform(rolls: np.ndarray, H: int, W: int, random_crop: bool, max_pitch_shift: int, binarize_threshold: float,
                   mean: float, std: float, seed: int, offset: int):
    """rolls uint8 [B, Hs, Ws] -> (float32 [B, 1, H, W], per-sample (top, left, shift))."""
    B, Hs, Ws = rolls.shape
    out = np.zeros((B, 1, H, W), dtype=np.float32)
    params = []
    inv255, inv_std = np.float32(1.0 / 255.0), np.float32(1.0) / np.float32(std)
    for b in range(B):
        r = sample_draw(seed, offset + b)
        if random_crop:
            top, left = r[0] % (Hs - H + 1), r[1] % (Ws - W + 1)
        else:
            top, left = center_offset(Hs, H), center_offset(Ws, W)
        shift = (r[2] % (2 * max_pitch_shift + 1)) - max_pitch_shift if max_pitch_shift > 0 else 0
        win = rolls[b, top:top + H, left:left + W].astype(np.float32) * inv255
        moved = np.zeros_like(win)
        lo, hi = max(0, -shift), min(H, H - shift)   # rows h with 0 <= h + shift < H
        moved[lo:hi] = win[lo + shift:hi + shift]
        if binarize_threshold >= 0.0:
            moved = (moved > np.float32(binarize_threshold)).astype(np.float32)
        out[b, 0] = (moved - np.float32(mean)) * inv_std
        params.append((top, left, shift))
    return out, params

"""Device-side piano-roll transforms (north_star item 3).

Mirror of the reference's ``midi_autoencoder/data_transformations.py:14-119``
(``get_transform``) for uint8 piano-rolls that are ALREADY resident in HBM:
instead of per-sample torchvision pipelines in DataLoader workers
(Resize -> Random/CenterCrop -> ToImage -> ToDtype(float32, scale=True) ->
Normalize(mean, std) [-> Grayscale]) one CUDA kernel crops, scales by 1/255 and
normalises a whole batch, and can additionally binarise and pitch-transpose
(the two augmentations the north star adds; the reference has neither).

Same registry and argument conventions as the reference: ``NORMALIZATION``
(:4-9), ``get_transform(transform_type, image_size, args)`` returning
``(train_transform, test_transform)``, ``NotImplementedError`` for unknown
types (:117).  ``Resize`` is the identity for rolls whose short side already
equals ``image_size`` (128 pitches) and is otherwise not supported on device.
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops

NORMALIZATION = {
    "mnist": [(0.5,), (1.0,)],
    "vae-lines": [(0.5,), (1.0,)],
    "vae-lines-large": [(0.5,), (1.0,)],
    # {0,1} targets for the BCE loss (the reference's Normalize(0.5, 1.0) yields [-0.5, 0.5],
    # which torch >= 2.x binary_cross_entropy rejects)
    "piano-roll": [(0.0,), (1.0,)],
}

VALID_TRANSFORMS = ["mnist", "vae-lines", "vae-lines-large", "piano-roll"]


class RollTransform:
    """Callable ``uint8 [B, Hs, Ws] (cuda) -> float32 [B, 1, H, W]``."""

    def __init__(self, height: int, width: int, random_crop: bool, mean: float, std: float,
                 binarize_threshold: Optional[float] = None, max_pitch_shift: int = 0, seed: int = 0):
        self.height, self.width = height, width
        self.random_crop = random_crop
        self.mean, self.std = float(mean), float(std)
        self.binarize_threshold = -1.0 if binarize_threshold is None else float(binarize_threshold)
        self.max_pitch_shift = int(max_pitch_shift)
        self.seed = int(seed)
        self.offset = 0  # Philox counter: advances by the batch size per call

    def manual_seed(self, seed: int, rank: int = 0) -> None:
        self.seed, self.offset = int(seed) + int(rank), 0

    def __call__(self, rolls: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if rolls.dim() == 4 and rolls.shape[1] == 1:
            rolls = rolls[:, 0]
        if rolls.shape[1] < self.height or rolls.shape[2] < self.width:
            raise RuntimeError(f"rolls {tuple(rolls.shape[1:])} smaller than the crop {(self.height, self.width)}")
        out = ops.roll_transform(
            rolls.contiguous(), self.height, self.width, 1 if self.random_crop else 0,
            self.max_pitch_shift if self.random_crop else 0, self.binarize_threshold, self.mean, self.std,
            self.seed, self.offset, out=out,
        )
        self.offset += rolls.shape[0]
        return out


def get_transform(transform_type: str = "noaug", image_size=32, args=None) -> Tuple[RollTransform, RollTransform]:
    """(train_transform, test_transform), as data_transformations.py:14.

    ``image_size``: int (square) or (H, W).  ``args`` accepts the reference's keys
    (``normalization``, ``mean``, ``std``) plus ``binarize_threshold``,
    ``max_pitch_shift`` and ``seed``."""
    if args is None:
        args = {}
    mean, std = NORMALIZATION[args.get("normalization", "mnist")]
    if "mean" in args:
        mean = args["mean"]
    if "std" in args:
        std = args["std"]
    mean = float(mean[0] if isinstance(mean, (tuple, list)) else mean)
    std = float(std[0] if isinstance(std, (tuple, list)) else std)
    h, w = (image_size, image_size) if isinstance(image_size, int) else tuple(image_size)
    thr = args.get("binarize_threshold")
    seed = args.get("seed", 0)
    if transform_type in ("noaug", "midi"):
        # random crop for training, centre crop for testing (data_transformations.py:31-32,41; :54,64)
        shift = args.get("max_pitch_shift", 0) if transform_type == "midi" else 0
        train = RollTransform(h, w, True, mean, std, thr, shift, seed)
        test = RollTransform(h, w, False, mean, std, thr, 0, seed)
    elif transform_type == "digits":
        # centre crop for both (data_transformations.py:95-103,108-114)
        train = RollTransform(h, w, False, mean, std, thr, 0, seed)
        test = RollTransform(h, w, False, mean, std, thr, 0, seed)
    else:
        raise NotImplementedError
    return (train, test)

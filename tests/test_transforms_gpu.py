"""Device-side piano-roll transform (north_star item 3) against
  (1) the reference's OWN torchvision pipelines (data_transformations.py:23-72, imported from baseline/_ref and
      run on the CPU) -- crop / scale / normalise semantics, bit-exact;
  (2) oracle/roll_transform_oracle.py -- the numpy restatement that also pins the Philox-derived crop offsets and
      the two new augmentations (binarise, pitch transpose), bit-exact.
"""
import numpy as np
import pytest
import torch

from oracle import reference_driver as rd
from oracle import roll_transform_oracle as ro

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not rd.available(), reason="baseline/_ref (the installed reference) is not present")


@pytest.fixture(scope="module")
def tv():
    import torch_vae_b200 as m

    m.load_library()
    return m


def make_rolls(B, Hs, Ws, seed=0, density=0.1, grey=False):
    g = torch.Generator().manual_seed(seed)
    if grey:
        return torch.randint(0, 256, (B, Hs, Ws), generator=g, dtype=torch.int32).to(torch.uint8)
    return ((torch.rand(B, Hs, Ws, generator=g) < density) * 255).to(torch.uint8)


@needs_ref
@pytest.mark.parametrize("kind", ["midi", "noaug"])
def test_center_crop_stack_equals_reference_torchvision_pipeline(tv, kind):
    """test_transform of data_transformations.py:36-45 / :62-71, sample by sample, on grey-level rolls."""
    from torchvision import tv_tensors

    ref_mod = rd.import_module("data_transformations")
    _, ref_test = ref_mod.get_transform(kind, 128, {"normalization": "mnist"})
    _, got_test = tv.data_transformations.get_transform(kind, 128, {"normalization": "mnist"})
    for Ws in (128, 160, 161, 250):
        rolls = make_rolls(4, 128, Ws, seed=Ws, grey=True)
        got = got_test(rolls.cuda()).cpu()
        for b in range(rolls.shape[0]):
            want = ref_test(tv_tensors.Image(rolls[b][None]))
            assert want.shape == (1, 128, 128)
            assert torch.equal(got[b], torch.as_tensor(want)), (kind, Ws, b)


@needs_ref
def test_random_crop_stack_equals_reference_pipeline_at_the_same_window(tv):
    """train_transform of data_transformations.py:52-61: the window comes from a different RNG (device Philox
    instead of torch's CPU generator), so parity is: the kernel's output == the reference's own ToImage ->
    ToDtype(scale) -> Normalize -> Grayscale tail applied to the crop at the kernel's window, and the windows
    are uniform over the same range RandomCrop draws from (torch.randint(0, w - tw + 1))."""
    from torchvision import tv_tensors
    from torchvision.transforms.v2 import functional as TF

    ref_mod = rd.import_module("data_transformations")
    ref_train, _ = ref_mod.get_transform("midi", 128, {"normalization": "mnist"})
    names = [type(t).__name__ for t in ref_train.transforms]
    assert names == ["Resize", "RandomCrop", "ToImage", "ToDtype", "Normalize", "Grayscale"], names
    tail = ref_train.transforms[2:]
    got_train, _ = tv.data_transformations.get_transform("midi", 128, {"normalization": "mnist", "seed": 21})
    B, Ws = 256, 160
    rolls = make_rolls(B, 128, Ws, seed=5, grey=True)
    got = got_train(rolls.cuda()).cpu()
    want, prm = ro.roll_transform(rolls.numpy(), 128, 128, True, 0, -1.0, 0.5, 1.0, 21, 0)
    assert np.array_equal(got.numpy(), want)  # Philox-derived windows pinned by the oracle
    lefts = [p[1] for p in prm]
    for b in range(0, B, 8):
        img = TF.crop(tv_tensors.Image(rolls[b][None]), 0, lefts[b], 128, 128)
        for t in tail:
            img = t(img)
        assert torch.equal(got[b], torch.as_tensor(img)), b
    assert min(lefts) == 0 and max(lefts) == Ws - 128 and len(set(lefts)) >= 30  # 33 possible windows, 256 draws
    counts = np.bincount(lefts, minlength=33)
    assert counts.max() <= 24, counts  # mean 7.8 per window: no window is grossly over-represented
    # a second call advances the Philox counter (fresh windows); the same (seed, offset) reproduces
    got2 = got_train(rolls.cuda()).cpu()
    assert not torch.equal(got, got2)
    got_train.manual_seed(21)
    assert torch.equal(got_train(rolls.cuda()).cpu(), got)


@pytest.mark.parametrize("cfg", [
    # Hs, Ws, H, W, random_crop, max_shift, threshold, mean, std, seed, offset
    (128, 160, 128, 128, True, 6, 0.5, 0.0, 1.0, 11, 0),
    (128, 160, 128, 128, False, 6, 0.5, 0.0, 1.0, 11, 7),      # test transform: no transpose
    (140, 1100, 128, 1024, True, 12, -1.0, 0.5, 1.0, 3, 100),  # long-context rolls, grey levels, pitch crop too
    (128, 131, 128, 125, True, 3, 0.25, 0.1, 2.0, 5, 1),       # width not a multiple of 8: scalar path
])
def test_binarise_and_pitch_transpose_are_pinned_by_the_oracle(tv, cfg):
    Hs, Ws, H, W, rc, ms, thr, mean, std, seed, off = cfg
    rolls = make_rolls(6, Hs, Ws, seed=seed, grey=True)
    got = tv.ops.roll_transform(rolls.cuda(), H, W, 1 if rc else 0, ms if rc else 0, thr, mean, std, seed, off).cpu().numpy()
    want, prm = ro.roll_transform(rolls.numpy(), H, W, rc, ms if rc else 0, thr, mean, std, seed, off)
    assert np.array_equal(got, want)
    if rc and ms > 0:
        assert len({p[2] for p in prm}) > 1  # different samples, different transpositions
    if thr >= 0:
        assert set(np.unique(got)) <= {np.float32((0 - mean) / std), np.float32((1 - mean) / std)}

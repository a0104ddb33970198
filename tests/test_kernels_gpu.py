"""Per-kernel parity: every C-ABI entry point against a CPU fp64 restatement of
the same ATen op the reference dispatches to, fp32 and bf16.

Tolerances (norm-wise relative error, written here on purpose):
  fp32 kernels : 1e-5  (north_star bound for fp32)
  bf16 kernels : 6e-3 per op against fp64 math on the bf16-rounded inputs
                 (one bf16 output rounding is 2^-9 = 2e-3 per element)
"""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle.vae_oracle import rel_err

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-5, torch.bfloat16: 6e-3}
DTYPES = [torch.float32, torch.bfloat16]


@pytest.fixture(scope="module")
def tv():
    import torch_vae_b200 as m

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    m.load_library()
    return m


def dev(t):
    return t.cuda()


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def q(t, dtype):
    """Round to the compute dtype and come back to fp64 (what the kernel sees)."""
    return t.to(dtype).double()


CONV_CASES = [
    # B, H, W, Ci, Co, stride
    (2, 16, 16, 32, 64, 2),
    (3, 8, 24, 64, 128, 2),
    (2, 8, 8, 128, 256, 2),
    (2, 16, 16, 8, 16, 2),
    (2, 12, 20, 24, 40, 2),  # channel counts that are not powers of two
    (2, 32, 32, 1, 32, 2),  # enc.conv0 edge kernel
    (2, 16, 16, 3, 16, 2),  # multi-channel image through the generic kernel
    (2, 16, 24, 32, 1, 1),  # final conv edge kernel
    (1, 8, 8, 16, 16, 1),
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fprop_dgrad_wgrad(tv, case, dtype):
    B, H, W, Ci, Co, s = case
    ops = tv.ops
    x = rnd(B, Ci, H, W, seed=1)
    w = rnd(Co, Ci, 3, 3, seed=2, scale=1.0 / math.sqrt(9 * Ci))
    b = rnd(Co, seed=3, scale=0.1)
    in_dtype = torch.float32 if Ci in (1, 3) else dtype  # the image itself is always fp32
    xq, wq = q(x, in_dtype), q(w, dtype)
    y_ref = F.conv2d(xq, wq, b.double(), stride=s, padding=1)

    wp, wt = ops.pack_conv_weight(dev(w), dtype)
    stats = torch.empty(2 * Co, dtype=torch.float64, device="cuda")
    x_dev = dev(nhwc(x)).to(in_dtype)
    y = ops.conv_fprop(x_dev, wp, dev(b), s, dtype, stats if Co > 1 else None, Ci=Ci)
    assert y.shape == (B, H // s, W // s, Co)
    assert rel_err(nchw(y.float().cpu()), y_ref) < TOL[dtype]
    if Co > 1:
        st = stats.cpu()
        assert rel_err(st[:Co], y_ref.sum(dim=(0, 2, 3))) < 1e-3 + TOL[dtype]
        assert rel_err(st[Co:], (y_ref**2).sum(dim=(0, 2, 3))) < 1e-3 + TOL[dtype]

    # data gradient == conv_transpose2d of dy with the same weight
    dy = rnd(B, Co, H // s, W // s, seed=4)
    dyq = q(dy, dtype)
    dx_ref = F.conv_transpose2d(dyq, wq, None, stride=s, padding=1, output_padding=s - 1)
    bias_t = rnd(Ci, seed=5, scale=0.1)
    stats2 = torch.empty(2 * Ci, dtype=torch.float64, device="cuda")
    dx = ops.conv_dgrad(dev(nhwc(dy)).to(dtype), wt, dev(bias_t), s, stats2)
    dx_ref_b = dx_ref + bias_t.double().view(1, -1, 1, 1)
    assert dx.shape == (B, H, W, Ci)
    assert rel_err(nchw(dx.float().cpu()), dx_ref_b) < TOL[dtype]
    assert rel_err(stats2.cpu()[:Ci], dx_ref_b.sum(dim=(0, 2, 3))) < 1e-3 + TOL[dtype]

    # weight gradient
    xr = xq.clone().requires_grad_(True)
    wr = wq.clone().requires_grad_(True)
    F.conv2d(xr, wr, None, stride=s, padding=1).backward(dyq)
    dw = ops.conv_wgrad(x_dev, dev(nhwc(dy)).to(dtype), s, Ci, Co, H, W)
    assert dw.shape == (Co, Ci, 3, 3) and dw.dtype == torch.float32
    assert rel_err(dw.cpu(), wr.grad) < (1e-5 if dtype == torch.float32 else 1e-4)


@pytest.mark.parametrize("one_launch", [True, False], ids=["cooperative_fused_backward", "two_pass_backward"])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("C,n_pix_shape", [(32, (4, 16, 16)), (8, (2, 8, 8)), (24, (3, 4, 6)), (20, (2, 4, 4)), (128, (64, 16, 16)),
                                           (64, (96, 32, 32))])
def test_batchnorm_lrelu_forward_backward(tv, C, n_pix_shape, dtype, one_launch, monkeypatch):
    ops = tv.ops
    monkeypatch.setattr(ops, "FUSE_BN_BWD_PASSES", one_launch)   # both passes in one cooperative launch, or two kernels
    B, H, W = n_pix_shape
    y = rnd(B, C, H, W, seed=1, scale=2.0) + 0.3
    gamma, beta = rnd(C, seed=2) * 0.2 + 1.0, rnd(C, seed=3) * 0.1
    dact = rnd(B, C, H, W, seed=4)
    yq, dq = q(y, dtype), q(dact, dtype)
    rm, rv = torch.zeros(C), torch.ones(C)

    yr = yq.clone().requires_grad_(True)
    gr, br = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rm_ref, rv_ref = rm.double().clone(), rv.double().clone()
    a_ref = F.leaky_relu(F.batch_norm(yr, rm_ref, rv_ref, gr, br, True, 0.1, 1e-5), 0.01)
    a_ref.backward(dq)

    y_dev = dev(nhwc(y)).to(dtype)
    stats = torch.stack([yq.sum(dim=(0, 2, 3)), (yq**2).sum(dim=(0, 2, 3))]).flatten().cuda()
    rm_d, rv_d = dev(rm), dev(rv)
    vec = ops.bn_finalize(stats, dev(gamma), dev(beta), rm_d, rv_d, B * H * W, True)
    assert rel_err(rm_d.cpu(), rm_ref) < 1e-5 and rel_err(rv_d.cpu(), rv_ref) < 1e-5
    act = ops.bn_apply_lrelu(y_dev, vec, dev(beta))
    assert rel_err(nchw(act.float().cpu()), a_ref) < TOL[dtype]
    dy, dgamma, dbeta, dbias = ops.bn_lrelu_backward(dev(nhwc(dact)).to(dtype), y_dev, vec, dev(gamma), dev(beta), True)
    assert rel_err(nchw(dy.float().cpu()), yr.grad) < max(TOL[dtype], 2e-5)
    assert rel_err(dgamma.cpu(), gr.grad) < 2e-5 and rel_err(dbeta.cpu(), br.grad) < 2e-5
    assert float(dbias.abs().max()) < 1e-3 * float(yr.grad.abs().sum())

    # eval mode: running statistics
    vec_e = ops.bn_finalize(None, dev(gamma), dev(beta), rm_d, rv_d, B * H * W, False)
    a_eval = F.leaky_relu(F.batch_norm(yq, rm_ref, rv_ref, gamma.double(), beta.double(), False, 0.1, 1e-5), 0.01)
    assert rel_err(nchw(ops.bn_apply_lrelu(y_dev, vec_e, dev(beta)).float().cpu()), a_eval) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
def test_layout_transposes(tv, dtype):
    ops = tv.ops
    x = rnd(3, 40, 6, 10, seed=1).to(dtype)
    flat = ops.nhwc_to_nchw(dev(nhwc(x)))
    assert torch.equal(flat.cpu(), x.flatten(1))
    back = ops.nchw_to_nhwc(dev(x.flatten(1)), 40, 6, 10)
    assert torch.equal(back.cpu(), nhwc(x))
    assert torch.equal(ops.nhwc_to_nchw(dev(nhwc(x)), torch.float32).cpu(), x.flatten(1).float())


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,K,N", [(8, 1024, 10), (64, 16384, 10), (5, 300, 33), (16, 10, 1024)])
def test_linear_forward_backward(tv, dtype, B, K, N):
    ops = tv.ops
    x, w, b = rnd(B, K, seed=1), rnd(N, K, seed=2, scale=1 / math.sqrt(K)), rnd(N, seed=3)
    dout = rnd(B, N, seed=4)
    xq = q(x, dtype)
    y = ops.linear_fwd(dev(x).to(dtype), dev(w), dev(b), torch.float32)
    assert rel_err(y.cpu(), F.linear(xq, w.double(), b.double())) < 1e-5
    y2 = ops.linear_fwd(dev(x), dev(w), dev(b), dtype)  # fp32 in, compute-dtype out (decoder_input)
    assert rel_err(y2.float().cpu(), F.linear(x.double(), w.double(), b.double())) < TOL[dtype]
    dx = ops.linear_bwd_input(dev(dout), dev(w), out_dtype=dtype)
    assert rel_err(dx.float().cpu(), dout.double() @ w.double()) < TOL[dtype]
    ops.linear_bwd_input(dev(dout), dev(w), out=dx, accumulate=True)
    assert rel_err(dx.float().cpu(), 2 * (dout.double() @ w.double())) < 2 * TOL[dtype]
    dw = ops.linear_bwd_weight(dev(dout), dev(x).to(dtype))
    assert rel_err(dw.cpu(), dout.double().t() @ xq) < 1e-5
    assert rel_err(ops.colsum(dev(dout)).cpu(), dout.double().sum(0)) < 1e-5
    assert rel_err(ops.colsum(dev(x).to(dtype)).cpu(), xq.sum(0)) < 1e-5


def test_reparam_injected_and_philox(tv):
    ops = tv.ops
    mu, lv, eps = rnd(7, 10, seed=1), rnd(7, 10, seed=2), rnd(7, 10, seed=3)
    z, e = ops.reparam_fwd(dev(mu), dev(lv), dev(eps))
    assert rel_err(z.cpu(), eps.double() * torch.exp(0.5 * lv.double()) + mu.double()) < 1e-6
    dz = rnd(7, 10, seed=4)
    dmu, dlv = ops.reparam_bwd(dev(dz), dev(eps), dev(lv))
    assert torch.equal(dmu.cpu(), dz)
    assert rel_err(dlv.cpu(), dz.double() * eps.double() * 0.5 * torch.exp(0.5 * lv.double())) < 1e-6

    n = 1 << 20
    a = ops.philox_randn(n, 1234, 0, "cuda")
    b = ops.philox_randn(n, 1234, 0, "cuda")
    c = ops.philox_randn(n, 1234, n // 4, "cuda")
    d = ops.philox_randn(n, 1235, 0, "cuda")
    assert torch.equal(a, b), "(seed, offset) must reproduce the stream"
    assert not torch.equal(a, c) and not torch.equal(a, d)
    assert abs(float(a.mean())) < 5e-3 and abs(float(a.std()) - 1.0) < 5e-3
    assert abs(float((a**4).mean()) - 3.0) < 0.05  # kurtosis of N(0,1)
    assert abs(float((a * c).mean())) < 5e-3  # consecutive blocks uncorrelated
    # device-side offset (CUDA-graph replays) equals the host-side offset
    off = torch.tensor([n // 4], dtype=torch.int64, device="cuda")
    zero = torch.zeros(n, device="cuda")
    z_off, _ = ops.reparam_fwd(zero, zero, None, seed=1234, offset=0, dev_offset=off)
    assert torch.equal(z_off, c)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("kldw", [1.0, 0.00025])
def test_fused_elbo_matches_reference_semantics(tv, dtype, kldw):
    ops = tv.ops
    B, L, H, W = 6, 10, 16, 24
    logits = rnd(B, 1, H, W, seed=1, scale=3.0)
    logits[0, 0, 0, :4] = torch.tensor([25.0, -25.0, 110.0, -110.0])  # saturation / clamp edge
    x = (torch.rand(B, 1, H, W, generator=torch.Generator().manual_seed(2)) < 0.3).float()
    x[0, 0, 0, :4] = torch.tensor([0.0, 1.0, 0.0, 1.0])
    mu, lv = rnd(B, L, seed=3), rnd(B, L, seed=4) * 0.5
    lq = logits.to(dtype).float()

    # reference semantics in fp32 ATen on the CPU: sigmoid -> binary_cross_entropy (+ KL)
    lr = lq.clone().requires_grad_(True)
    mr, vr = mu.clone().requires_grad_(True), lv.clone().requires_grad_(True)
    recon = F.binary_cross_entropy(torch.sigmoid(lr), x)
    kld = -0.5 * torch.mean(torch.sum(1 + vr - mr**2 - torch.exp(vr), dim=-1))
    loss = recon + kldw * kld
    (loss * 1.7).backward()

    ld = dev(logits).to(dtype)
    losses = ops.elbo_fwd(ld, dev(x), dev(mu), dev(lv), kldw, False).cpu()
    assert abs(float(losses[0]) - float(loss)) <= 2e-6 * abs(float(loss)) + 1e-7
    assert abs(float(losses[1]) - float(recon)) <= 2e-6 * abs(float(recon)) + 1e-7
    assert abs(float(losses[2]) + float(kld)) <= 2e-6 * abs(float(kld)) + 1e-7  # sign-flipped
    g = torch.tensor(1.7, device="cuda")
    dl, dmu, dlv = ops.elbo_bwd(ld, dev(x), dev(mu), dev(lv), g, kldw, False)
    assert rel_err(dl.float().cpu(), lr.grad) < TOL[dtype]
    assert rel_err(dmu.cpu(), mr.grad) < 1e-5 and rel_err(dlv.cpu(), vr.grad) < 1e-5

    # probabilities path (hand-built ModelOutput): BCE on p directly
    p = torch.sigmoid(lq)
    pr = p.clone().requires_grad_(True)
    (F.binary_cross_entropy(pr, x) * 1.7).backward()
    lp = ops.elbo_fwd(dev(p), dev(x), dev(mu), dev(lv), kldw, True).cpu()
    assert abs(float(lp[1]) - float(recon)) <= 2e-6 * abs(float(recon)) + 1e-7
    dp, _, _ = ops.elbo_bwd(dev(p), dev(x), dev(mu), dev(lv), g, kldw, True)
    assert rel_err(dp.cpu(), pr.grad) < 1e-5


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(6, 10, 16, 24), (3, 7, 5, 9), (64, 10, 128, 128)])
def test_one_pass_elbo_equals_forward_plus_backward(tv, dtype, shape):
    """vae_elbo_fused == vae_elbo_fwd + vae_elbo_bwd (the reference-semantics pair tested above),
    including the clamp / saturation edge; vae_elbo_scale_grads applies the upstream gradient."""
    ops = tv.ops
    B, L, H, W = shape
    logits = rnd(B, 1, H, W, seed=1, scale=3.0)
    logits[0, 0, 0, :4] = torch.tensor([25.0, -25.0, 110.0, -110.0])
    logits[0, 0, 0, 4:8] = torch.tensor([88.0, -88.0, 95.0, -95.0])  # subnormal sigmoid: the MUFU path must step aside
    x = (torch.rand(B, 1, H, W, generator=torch.Generator().manual_seed(2)) < 0.3).float()
    x[0, 0, 0, :8] = torch.tensor([0.0, 1.0, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0])
    mu, lv = rnd(B, L, seed=3), rnd(B, L, seed=4) * 0.5
    ld, xd, md, vd = dev(logits).to(dtype), dev(x), dev(mu), dev(lv)
    one = torch.tensor(1.0, device="cuda")
    ref_losses = ops.elbo_fwd(ld, xd, md, vd, 0.5, False)
    ref_dl, ref_dmu, ref_dlv = ops.elbo_bwd(ld, xd, md, vd, one, 0.5, False)
    losses, dl, dmu, dlv = ops.elbo_fused(ld, xd, md, vd, 0.5, False)
    ltol = 2e-6 if dtype == torch.float32 else 5e-5   # bf16 path: MUFU log / exp
    for i in range(3):
        assert abs(float(losses[i]) - float(ref_losses[i])) <= ltol * abs(float(ref_losses[i])) + 1e-7
    gtol = 1e-6 if dtype == torch.float32 else 2e-3   # one bf16 rounding of nearly equal values
    assert rel_err(dl.float().cpu(), ref_dl.float().cpu().double()) <= gtol
    assert torch.equal(dl[0, 0, 0, 2:4].float().cpu(), ref_dl[0, 0, 0, 2:4].float().cpu())  # saturated: exactly 0 / clamp
    assert rel_err(dmu.cpu(), ref_dmu.cpu().double()) <= 1e-6 and rel_err(dlv.cpu(), ref_dlv.cpu().double()) <= 1e-6
    # upstream gradient: 1 leaves the buffers untouched, 1.7 scales them in place
    keep = dl.clone()
    ops.elbo_scale_grads(one, dl, dmu, dlv)
    assert torch.equal(dl, keep)
    ops.elbo_scale_grads(torch.tensor(1.7, device="cuda"), dl, dmu, dlv)
    ref_dl, ref_dmu, ref_dlv = ops.elbo_bwd(ld, xd, md, vd, torch.tensor(1.7, device="cuda"), 0.5, False)
    assert rel_err(dl.float().cpu(), ref_dl.float().cpu().double()) <= max(gtol, 4e-3 if dtype != torch.float32 else 0)
    assert rel_err(dmu.cpu(), ref_dmu.cpu().double()) <= 1e-6
    # probabilities as input (hand-built ModelOutput)
    pd = torch.sigmoid(ld.float())
    lp, dp, _, _ = ops.elbo_fused(pd, xd, md, vd, 0.5, True)
    rp = ops.elbo_fwd(pd, xd, md, vd, 0.5, True)
    rdp, _, _ = ops.elbo_bwd(pd, xd, md, vd, one, 0.5, True)
    assert abs(float(lp[0]) - float(rp[0])) <= 2e-6 * abs(float(rp[0])) + 1e-7
    assert rel_err(dp.cpu(), rdp.cpu().double()) <= 1e-6


def test_elbo_autograd_second_backward_and_scaled_grad(tv):
    """ElboFn: the cached one-pass gradients serve the first backward (any upstream gradient);
    a second backward through a retained graph recomputes them."""
    from torch_vae_b200 import engine
    B, L, H, W = 4, 10, 16, 16
    logits = dev(rnd(B, 1, H, W, seed=5, scale=2.0)).requires_grad_(True)
    x = dev((torch.rand(B, 1, H, W, generator=torch.Generator().manual_seed(6)) < 0.3).float())
    mu, lv = dev(rnd(B, L, seed=7)).requires_grad_(True), dev(rnd(B, L, seed=8) * 0.5).requires_grad_(True)
    loss, _ = engine.ElboFn.apply(logits, x, mu, lv, 0.7, False)
    (loss * 3.0).backward(retain_graph=True)
    g1 = [t.grad.clone() for t in (logits, mu, lv)]
    for t in (logits, mu, lv):
        t.grad = None
    (loss * 3.0).backward()
    for a, t in zip(g1, (logits, mu, lv)):
        assert rel_err(t.grad.cpu(), a.cpu().double()) <= 1e-6
    lr = logits.detach().cpu().clone().requires_grad_(True)
    mr, vr = mu.detach().cpu().clone().requires_grad_(True), lv.detach().cpu().clone().requires_grad_(True)
    ref = F.binary_cross_entropy(torch.sigmoid(lr), x.cpu()) + 0.7 * (-0.5 * torch.mean(torch.sum(1 + vr - mr**2 - torch.exp(vr), dim=-1)))
    (ref * 3.0).backward()
    assert rel_err(g1[0].cpu(), lr.grad) <= 1e-5 and rel_err(g1[1].cpu(), mr.grad) <= 1e-5 and rel_err(g1[2].cpu(), vr.grad) <= 1e-5
    with torch.no_grad():
        l2, _ = engine.ElboFn.apply(logits, x, mu, lv, 0.7, False)
    assert abs(float(l2) - float(ref)) <= 1e-5 * abs(float(ref))


def test_sigmoid_and_backward(tv):
    ops = tv.ops
    l = rnd(4, 1, 8, 8, seed=1, scale=4.0)
    p = ops.sigmoid(dev(l))
    assert rel_err(p.cpu(), torch.sigmoid(l.double())) < 1e-6
    dp = rnd(4, 1, 8, 8, seed=2)
    dl = ops.sigmoid_bwd(dev(dp), p)
    pd = torch.sigmoid(l.double())
    assert rel_err(dl.cpu(), dp.double() * pd * (1 - pd)) < 1e-5


def test_roll_transform_matches_midi_stack(tv):
    """data_transformations.py:48-72: crop -> /255 -> Normalize(0.5, 1.0)."""
    ops = tv.ops
    g = torch.Generator().manual_seed(0)
    rolls = ((torch.rand(5, 128, 160, generator=g) < 0.1) * 255).to(torch.uint8)
    out = ops.roll_transform(dev(rolls), 128, 128, 0, 0, -1.0, 0.5, 1.0, 0, 0).cpu()
    ref = rolls[:, :, 16:144].float().div(255).sub(0.5).unsqueeze(1)  # CenterCrop: left = round(32/2)
    assert torch.equal(out, ref)
    # random crop: every sample is SOME window of its source; reproducible in (seed, offset)
    o1 = ops.roll_transform(dev(rolls), 128, 128, 1, 0, 0.5, 0.0, 1.0, 7, 3).cpu()
    o2 = ops.roll_transform(dev(rolls), 128, 128, 1, 0, 0.5, 0.0, 1.0, 7, 3).cpu()
    assert torch.equal(o1, o2)
    for b in range(5):
        wins = rolls[b].unfold(1, 128, 1).permute(1, 0, 2)  # [33,128,128]
        assert any(torch.equal(o1[b, 0], (w > 127).float()) for w in wins)
    # pitch transpose keeps the note count inside the window or drops rows off the edge
    o3 = ops.roll_transform(dev(rolls), 128, 128, 0, 6, 0.5, 0.0, 1.0, 11, 0).cpu()
    assert set(o3.unique().tolist()) <= {0.0, 1.0}
    assert float(o3.sum()) <= float((rolls[:, :, 16:144] > 127).sum()) + 1e-3


# ---------------------------------------------------------------------------
# fused output block (models.py:76-82): BN apply + LeakyReLU + Conv2d(C->1) + Sigmoid
# forward, and its two-pass backward, against autograd on the same fp64 graph
# ---------------------------------------------------------------------------
FINAL_CASES = [
    # B, H, W, C
    (3, 16, 32, 32),
    (2, 32, 16, 8),
    (2, 8, 64, 16),
    (1, 24, 36, 64),  # partial tiles in both directions
    (2, 128, 128, 32),
    (3, 64, 128, 32),   # tcgen05 path (bf16): two 32-row bands per image
    (5, 48, 128, 32),   # tcgen05 path: 16-row bands
    (2, 32, 256, 32),   # tcgen05 path, rows of several 128-pixel tiles (forward: overlapping halo tiles of 126 outputs)
    (2, 16, 128, 64),   # tcgen05 path, two groups of 32 channels (partial logits travel in the fp32 probs buffer)
    (2, 32, 384, 64),   # both
    (1, 128, 1024, 64), # the output block of the wide / long configuration (one image)
]


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", FINAL_CASES)
def test_final_block_forward_backward(tv, case, dtype, training):
    B, H, W, C = case
    ops = tv.ops
    assert ops.final_block_supported(B, H, W, C)
    y = rnd(B, C, H, W, seed=11) * 1.5 + 0.3
    gamma = rnd(C, seed=12, scale=0.3) + 1.0
    beta = rnd(C, seed=13, scale=0.2)
    # bf16-representable weights: the tcgen05 kernels feed the tensor cores bf16 weight images, the
    # CUDA-core kernels read the fp32 master; with representable values both see the same numbers
    w_last = rnd(1, C, 3, 3, seed=14, scale=1.0 / math.sqrt(9 * C)).bfloat16().float()
    b_last = rnd(1, seed=15, scale=0.1)
    dlog = rnd(B, 1, H, W, seed=16, scale=1.0 / (B * H * W))
    run_mean, run_var = rnd(C, seed=17, scale=0.1) + 0.3, torch.rand(C, generator=torch.Generator().manual_seed(18)) + 1.5

    yq = q(y, dtype).requires_grad_(True)
    g64, b64 = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    w64, bl64 = w_last.double().requires_grad_(True), b_last.double().requires_grad_(True)
    u = F.batch_norm(yq, run_mean.double().clone(), run_var.double().clone(), g64, b64, training=training, eps=1e-5)
    logits_ref = F.conv2d(F.leaky_relu(u, 0.01), w64, bl64, padding=1)
    dlq = q(dlog, dtype)
    logits_ref.backward(dlq)

    y_d = dev(nhwc(y).to(dtype))
    n = B * H * W
    if training:
        yy = y_d.double()
        stats = torch.cat([yy.sum((0, 1, 2)), (yy * yy).sum((0, 1, 2))])
    else:
        stats = None
    rm, rv = dev(run_mean.clone()), dev(run_var.clone())
    vec = ops.bn_finalize(stats, dev(gamma), dev(beta), rm, rv, n, training)
    logits, probs = ops.final_conv_fwd(y_d, vec, dev(beta), dev(w_last), dev(b_last))
    tol = TOL[dtype]
    assert logits.shape == (B, 1, H, W) and probs.dtype == torch.float32
    assert rel_err(logits.float(), logits_ref.detach()) <= tol
    assert rel_err(probs, torch.sigmoid(logits.double())) <= 1e-6
    if ops.final_tc_supported(y_d):  # the CUDA-core kernels of the same library agree
        old = ops.set_final_tc(False)
        try:
            logits2, _ = ops.final_conv_fwd(y_d, vec, dev(beta), dev(w_last), dev(b_last))
        finally:
            ops.set_final_tc(old)
        assert rel_err(logits.float(), logits2.float()) <= tol

    dy, dgamma, dbeta, dbias, dw, db = ops.final_block_backward(
        dev(dlog.to(dtype)), y_d, vec, dev(gamma), dev(beta), dev(w_last), training)
    if ops.final_tc_supported(y_d):  # the CUDA-core kernels of the same library agree
        old = ops.set_final_tc(False)
        try:
            ref2 = ops.final_block_backward(dev(dlog.to(dtype)), y_d, vec, dev(gamma), dev(beta), dev(w_last), training)
        finally:
            ops.set_final_tc(old)
        assert rel_err(dy.float(), ref2[0].float()) <= tol
        # dw: the tensor-core path rounds the normalised activation to bf16 before the pixel sum
        for got, want, lim in zip((dgamma, dbeta, dw, db), (ref2[1], ref2[2], ref2[4], ref2[5]), (2e-3, 2e-3, 6e-3, 2e-3)):
            assert rel_err(got, want) <= lim
    # gradients: fp32 arithmetic on (for bf16) rounded operands, one output rounding for dy
    assert rel_err(nchw(dy.float()), yq.grad) <= tol
    gtol = 1e-5 if dtype == torch.float32 else 2e-3
    assert rel_err(dgamma, g64.grad) <= gtol
    assert rel_err(dbeta, b64.grad) <= gtol
    assert rel_err(dw, w64.grad) <= (6e-3 if ops.final_tc_supported(y_d) else gtol)
    assert rel_err(db, bl64.grad) <= gtol
    # bias of the conv feeding a train-mode BatchNorm: exactly zero in exact arithmetic
    scale_ref = float(yq.grad.abs().sum())
    if training:
        assert float(dbias.abs().max()) <= 1e-3 * scale_ref
    else:
        assert rel_err(dbias, yq.grad.sum((0, 2, 3))) <= max(gtol, 1e-4)


# ---------------------------------------------------------------------------
# skinny GEMMs of the latent heads (heads.cu)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(7, 1024, 10), (64, 16384, 10), (5, 4096, 6), (33, 2048, 16), (3, 264, 1)])
def test_skinny_head_gemms(tv, shape, dtype):
    ops = tv.ops
    M, Fd, L = shape
    assert ops.heads_supported(Fd, L)
    a = rnd(M, Fd, seed=21)
    w_lf = rnd(L, Fd, seed=22, scale=1.0 / math.sqrt(Fd))   # fc_mu / fc_var layout
    w_fl = rnd(Fd, L, seed=23, scale=1.0 / math.sqrt(L))    # decoder_input layout
    b_l, b_f = rnd(L, seed=24, scale=0.1), rnd(Fd, seed=25, scale=0.1)
    d0, d1 = rnd(M, L, seed=26), rnd(M, L, seed=27)
    aq = q(a, dtype)
    tol = TOL[dtype]
    # reduce: [L,F] layout (forward heads) and [F,L] layout (dz)
    out = ops.skinny_reduce(dev(a.to(dtype)), dev(w_lf), dev(b_l), Fd, 1, L)
    assert rel_err(out, aq @ w_lf.double().t() + b_l.double()) <= 1e-5
    out = ops.skinny_reduce(dev(a.to(dtype)), dev(w_fl), None, 1, L, L)
    assert rel_err(out, aq @ w_fl.double()) <= 1e-5
    # both heads in one pass over the features (fc_mu and fc_var share their input)
    w2, b2 = rnd(L, Fd, seed=28, scale=1.0 / math.sqrt(Fd)), rnd(L, seed=29, scale=0.1)
    o0, o1 = ops.skinny_reduce_pair(dev(a.to(dtype)), dev(w_lf), dev(b_l), dev(w2), dev(b2), Fd, 1, L)
    assert rel_err(o0, aq @ w_lf.double().t() + b_l.double()) <= 1e-5
    assert rel_err(o1, aq @ w2.double().t() + b2.double()) <= 1e-5
    # weight gradients in both layouts, with the column sums
    g, cs = ops.skinny_wgrad(dev(d0), dev(a.to(dtype)), (L, Fd), Fd, 1, False)
    assert cs is None and rel_err(g, d0.double().t() @ aq) <= 1e-5
    g, cs = ops.skinny_wgrad(dev(d0), dev(a.to(dtype)), (Fd, L), 1, L, True)
    assert rel_err(g, aq.t() @ d0.double()) <= 1e-5
    assert rel_err(cs, aq.sum(0)) <= 1e-5
    # expand: decoder_input forward and the summed input gradient of the two heads
    out = ops.skinny_expand(dev(d0), dev(w_fl), None, None, dev(b_f), 1, L, Fd, dtype)
    assert rel_err(out.float(), d0.double() @ w_fl.double().t() + b_f.double()) <= tol
    out = ops.skinny_expand(dev(d0), dev(w_lf), dev(d1), dev(w_lf * 0.5), None, Fd, 1, Fd, dtype)
    assert rel_err(out.float(), (d0.double() + 0.5 * d1.double()) @ w_lf.double()) <= tol


def test_fused_adamw_matches_torch(tv):
    """FusedAdamW == torch.optim.AdamW (same groups, lrs, weight decay) over several steps,
    and its state_dict loads into the stock optimizer."""
    torch.manual_seed(3)
    shapes = [(32, 1, 3, 3), (32,), (64, 32, 3, 3), (64,), (7, 5)]
    ref_p = [torch.randn(s, device="cuda") for s in shapes]
    got_p = [p.clone() for p in ref_p]
    def groups(ps):
        return [{"params": ps[:2], "lr": 3e-3}, {"params": ps[2:], "lr": 1e-3, "weight_decay": 0.05}]
    ref_p = [p.requires_grad_(True) for p in ref_p]
    got_p = [p.requires_grad_(True) for p in got_p]
    ref = torch.optim.AdamW(groups(ref_p), lr=1e-3, weight_decay=0.01)
    got = tv.FusedAdamW(groups(got_p), lr=1e-3, weight_decay=0.01)
    for it in range(5):
        for a, b in zip(ref_p, got_p):
            g = torch.randn_like(a) * (0.1 + it)
            a.grad, b.grad = g.clone(), g.clone()
        ref.step()
        got.step()
    torch.cuda.synchronize()
    for a, b in zip(ref_p, got_p):
        assert rel_err(b, a) <= 2e-6
    for a, b in zip(ref_p, got_p):
        assert rel_err(got.state[b]["exp_avg_sq"], ref.state[a]["exp_avg_sq"]) <= 2e-6
        assert float(got.state[b]["step"]) == 5.0
    ref2 = torch.optim.AdamW(groups(got_p), lr=1e-3, weight_decay=0.01)
    ref2.load_state_dict(got.state_dict())  # checkpoint compatibility (train.py:445-460)


# ---------------------------------------------------------------------------
# tensor-core Linear kernels for a wide latent (linear_tc.cu): fc_mu / fc_var / decoder_input at L = 64 .. 256
# ---------------------------------------------------------------------------
LT_CASES = [(64, 256, 8192), (8, 256, 4096), (33, 64, 4160), (64, 128, 12288), (16, 192, 8192)]


@pytest.mark.parametrize("case", LT_CASES, ids=[f"M{m}-L{l}-F{f}" for m, l, f in LT_CASES])
def test_wide_latent_linear_kernels(tv, case):
    """Every routed call of ops.linear_* against fp64 matmuls of the bf16-rounded operands (the kernels round the
    fp32 master weights and the fp32 latent operands to bf16 when they form the tensor-core fragments)."""
    ops = tv.ops
    M, L, F_ = case
    assert ops.linear_tc_supported(M, L, F_)
    g = torch.Generator().manual_seed(5)
    bf = lambda t: t.bfloat16().double()
    pre = (torch.randn(M, F_, generator=g) * 0.5).bfloat16()
    w_lf = torch.randn(L, F_, generator=g) / math.sqrt(F_)
    w_fl = torch.randn(F_, L, generator=g) / math.sqrt(L)
    b_l, b_f = torch.randn(L, generator=g), torch.randn(F_, generator=g)
    d0, d1 = torch.randn(M, L, generator=g), torch.randn(M, L, generator=g)
    l0 = ops.LAUNCHES
    # fc_mu / fc_var forward
    mu = ops.linear_fwd(pre.cuda(), w_lf.cuda(), b_l.cuda(), torch.float32)
    assert rel_err(mu.cpu(), pre.double() @ bf(w_lf).t() + b_l.double()) <= 2e-5
    # decoder_input forward (bf16 output)
    dec = ops.linear_fwd(d0.cuda(), w_fl.cuda(), b_f.cuda(), torch.bfloat16)
    assert rel_err(dec.float().cpu(), bf(d0) @ bf(w_fl).t() + b_f.double()) <= 3e-3
    # input gradient of the two heads, one pass
    dpre = ops.linear_bwd_input_pair(d0.cuda(), w_lf.cuda(), d1.cuda(), (w_lf * 0.5 + 0.01).cuda(), torch.bfloat16)
    assert rel_err(dpre.float().cpu(), bf(d0) @ bf(w_lf) + bf(d1) @ bf(w_lf * 0.5 + 0.01)) <= 3e-3
    # latent gradient of decoder_input
    dz = ops.linear_bwd_input(pre.cuda(), w_fl.cuda(), out_dtype=torch.float32)
    assert rel_err(dz.cpu(), pre.double() @ bf(w_fl)) <= 2e-5
    # weight gradients in both layouts
    gw = ops.linear_bwd_weight(d0.cuda(), pre.cuda())
    assert gw.shape == (L, F_) and rel_err(gw.cpu(), bf(d0).t() @ pre.double()) <= 2e-5
    gw_in = ops.linear_bwd_weight(pre.cuda(), d0.cuda())
    assert gw_in.shape == (F_, L) and rel_err(gw_in.cpu(), pre.double().t() @ bf(d0)) <= 2e-5
    assert ops.LAUNCHES - l0 == 6   # one launch each: no packing pass, no second accumulation pass

"""tcgen05 / TMEM conv kernels against a CPU fp64 restatement (F.conv2d /
F.conv_transpose2d on the bf16-rounded operands) and against the CUDA-core
kernels of the same library.  Tolerance 6e-3 norm-wise (bf16 output rounding)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle.vae_oracle import rel_err

pytestmark = pytest.mark.gpu
TOL = 6e-3


@pytest.fixture(scope="module")
def ops():
    import torch_vae_b200 as tv

    tv.load_library()
    assert tv.ops.umma_available()
    return tv.ops


def rnd(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


CASES = [
    # B, H, W, Ci, Co, stride   (H, W = high-resolution side)
    (2, 16, 16, 32, 64, 2),
    (3, 8, 24, 64, 128, 2),
    (2, 8, 8, 128, 256, 2),
    (2, 32, 32, 32, 32, 2),
    (5, 16, 16, 64, 32, 2),
    (1, 8, 8, 32, 64, 1),
    (7, 48, 16, 96, 48, 2),  # not powers of two; M not a multiple of 128
    (64, 64, 64, 32, 64, 2),  # many tiles per CTA: exercises the persistent loop + TMEM double buffer
]


@pytest.mark.parametrize("case", CASES)
def test_umma_fprop(ops, case):
    B, H, W, Ci, Co, s = case
    x = rnd(B, Ci, H, W, seed=1)
    w = rnd(Co, Ci, 3, 3, seed=2, scale=1.0 / math.sqrt(9 * Ci))
    b = rnd(Co, seed=3, scale=0.1)
    xq, wq = x.bfloat16().double(), w.bfloat16().double()
    ref = F.conv2d(xq, wq, b.double(), stride=s, padding=1)
    wpk = ops.umma_pack_weight(w.cuda(), s, 0)
    assert wpk is not None
    stats = torch.empty(2 * Co, dtype=torch.float64, device="cuda")
    y = ops.conv_umma(nhwc(x).cuda().bfloat16(), wpk, b.cuda(), stats, Ci, Co, s, 0)
    torch.cuda.synchronize()
    assert rel_err(nchw(y.float().cpu()), ref) < TOL
    assert rel_err(stats.cpu()[:Co], ref.sum(dim=(0, 2, 3))) < 2e-3
    assert rel_err(stats.cpu()[Co:], (ref**2).sum(dim=(0, 2, 3))) < 2e-3


@pytest.mark.parametrize("case", CASES)
def test_umma_dgrad_is_conv_transpose(ops, case):
    B, H, W, Ci, Co, s = case
    dy = rnd(B, Co, H // s, W // s, seed=4)
    w = rnd(Co, Ci, 3, 3, seed=2, scale=1.0 / math.sqrt(9 * Co))
    b = rnd(Ci, seed=5, scale=0.1)
    ref = F.conv_transpose2d(dy.bfloat16().double(), w.bfloat16().double(), b.double(), stride=s, padding=1,
                             output_padding=s - 1)
    wpk = ops.umma_pack_weight(w.cuda(), s, 1)
    if Co % 32 != 0:  # the data gradient contracts over Co: not a tcgen05 shape -> CUDA-core kernel serves it
        assert wpk is None
        return
    assert wpk is not None
    stats = torch.empty(2 * Ci, dtype=torch.float64, device="cuda")
    dx = ops.conv_umma(nhwc(dy).cuda().bfloat16(), wpk, b.cuda(), stats, Ci, Co, s, 1)
    torch.cuda.synchronize()
    assert dx.shape == (B, H, W, Ci)
    assert rel_err(nchw(dx.float().cpu()), ref) < TOL
    assert rel_err(stats.cpu()[:Ci], ref.sum(dim=(0, 2, 3))) < 2e-3


def test_umma_fused_bn_lrelu_prologue(ops):
    B, H, W, Ci, Co, s = 4, 16, 16, 64, 128, 2
    x = rnd(B, Ci, H, W, seed=1)
    w = rnd(Co, Ci, 3, 3, seed=2, scale=1.0 / math.sqrt(9 * Ci))
    scale, mean, beta = rnd(Ci, seed=6) * 0.3 + 1.0, rnd(Ci, seed=8) * 0.5, rnd(Ci, seed=7) * 0.2
    xq = x.bfloat16().double()
    act = F.leaky_relu((xq - mean.double().view(1, -1, 1, 1)) * scale.double().view(1, -1, 1, 1) + beta.double().view(1, -1, 1, 1), 0.01)
    ref = F.conv2d(act.bfloat16().double(), w.bfloat16().double(), None, stride=s, padding=1)
    wpk = ops.umma_pack_weight(w.cuda(), s, 0)
    vec = torch.stack([scale, torch.zeros(Ci), mean, torch.ones(Ci)]).cuda()
    y = ops.conv_umma(nhwc(x).cuda().bfloat16(), wpk, None, None, Ci, Co, s, 0, pro=(vec, beta.cuda()))
    assert rel_err(nchw(y.float().cpu()), ref) < TOL


def test_umma_matches_cuda_core_kernels_at_layer_sizes(ops):
    """Same library, two kernel families, layer shapes of the default model."""
    for (B, H, W, Ci, Co) in [(8, 64, 64, 32, 64), (8, 16, 16, 128, 256), (4, 128, 128, 32, 32)]:
        x = rnd(B, H, W, Ci, seed=1).cuda().bfloat16()
        w = rnd(Co, Ci, 3, 3, seed=2, scale=1.0 / math.sqrt(9 * Ci)).cuda()
        wp, wt = ops.pack_conv_weight(w, torch.bfloat16)
        y_simt = ops.conv_fprop(x, wp, None, 2, torch.bfloat16)
        y_umma = ops.conv_umma(x, ops.umma_pack_weight(w, 2, 0), None, None, Ci, Co, 2, 0)
        assert rel_err(y_umma.float(), y_simt.float()) < TOL
        dy = rnd(B, H // 2, W // 2, Co, seed=3).cuda().bfloat16()
        d_simt = ops.conv_dgrad(dy, wt, None, 2)
        d_umma = ops.conv_umma(dy, ops.umma_pack_weight(w, 2, 1), None, None, Ci, Co, 2, 1)
        assert rel_err(d_umma.float(), d_simt.float()) < TOL


WG_CASES = [
    # B, H, W, Ci, Co, stride
    (2, 16, 16, 32, 64, 2),
    (3, 8, 24, 64, 128, 2),
    (4, 8, 8, 128, 256, 2),
    (2, 32, 32, 32, 32, 2),
    (5, 16, 16, 64, 32, 2),
    (1, 8, 8, 32, 64, 1),
    (7, 48, 16, 96, 48, 2),
    (32, 64, 64, 32, 64, 2),  # many k-blocks per CTA
    (2, 16, 16, 256, 64, 2),
]


@pytest.mark.parametrize("case", WG_CASES)
def test_umma_wgrad(ops, case):
    B, H, W, Ci, Co, s = case
    x = rnd(B, Ci, H, W, seed=1)
    dy = rnd(B, Co, H // s, W // s, seed=4)
    xr = x.bfloat16().double()
    wr = torch.zeros(Co, Ci, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(xr, wr, None, stride=s, padding=1).backward(dy.bfloat16().double())
    dw = ops.conv_weight_grad(nhwc(x).cuda().bfloat16(), nhwc(dy).cuda().bfloat16(), s, Ci, Co, H, W)
    torch.cuda.synchronize()
    assert dw.shape == (Co, Ci, 3, 3) and dw.dtype == torch.float32
    assert rel_err(dw.cpu(), wr.grad) < 1e-4
    # and the CUDA-core kernel of the same library agrees
    ops._L().vae_set_umma_wgrad(0)
    try:
        dw2 = ops.conv_weight_grad(nhwc(x).cuda().bfloat16(), nhwc(dy).cuda().bfloat16(), s, Ci, Co, H, W)
    finally:
        ops._L().vae_set_umma_wgrad(1)
    assert rel_err(dw, dw2) < 1e-4


WG_TMA_CASES = [
    # B, H, W, Ci, Co, stride : layer shapes of the default and the wide model
    (2, 128, 128, 32, 32, 2),
    (3, 64, 64, 32, 64, 2),
    (5, 32, 32, 64, 128, 2),
    (6, 16, 16, 128, 256, 2),
    (3, 16, 16, 256, 512, 2),
    (2, 16, 512, 64, 32, 2),
    (3, 16, 16, 32, 32, 1),
    (40, 64, 64, 32, 64, 2),    # many k-blocks per CTA: ring wrap-around
    (130, 16, 16, 128, 128, 2),
]


@pytest.mark.parametrize("case", WG_TMA_CASES)
def test_wgrad_tma(ops, case):
    B, H, W, Ci, Co, s = case
    assert ops._L().vae_wgrad_tma_eligible(B, H, W, Ci, Co, s)
    x = rnd(B, Ci, H, W, seed=1)
    dy = rnd(B, Co, H // s, W // s, seed=4)
    wr = torch.zeros(Co, Ci, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.bfloat16().double(), wr, None, stride=s, padding=1).backward(dy.bfloat16().double())
    xd, yd = nhwc(x).cuda().bfloat16(), nhwc(dy).cuda().bfloat16()
    dw = ops.conv_weight_grad(xd, yd, s, Ci, Co, H, W)
    torch.cuda.synchronize()
    assert dw.shape == (Co, Ci, 3, 3) and dw.dtype == torch.float32
    assert rel_err(dw.cpu(), wr.grad) < 1e-4
    old = ops._L().vae_set_wgrad_tma(0)  # the cp.async kernel of the same library agrees
    try:
        dw2 = ops.conv_weight_grad(xd, yd, s, Ci, Co, H, W)
    finally:
        ops._L().vae_set_wgrad_tma(old)
    assert rel_err(dw, dw2) < 1e-4


TMA_CASES = [
    # B, H, W, Ci, Co, stride: extents that decompose into whole output rows per 128-pixel tile
    (2, 128, 128, 32, 32, 2),   # Wj = 64: two rows per tile
    (3, 64, 64, 32, 64, 2),     # Wj = 32
    (5, 32, 32, 64, 128, 2),    # Wj = 16, 128B swizzle
    (6, 16, 16, 128, 256, 2),   # tile spans two images; two channel chunks per tap
    (2, 16, 512, 64, 32, 2),    # wide rolls: Wj = 256, tiles of 128 pixels inside a row
    (3, 16, 16, 32, 32, 1),     # stride 1
    (64, 64, 64, 32, 64, 2),    # many tiles per CTA
    (4, 16, 128, 256, 512, 2),  # wide model: 512 output channels = two 256-channel launches (UMMA N <= 256)
    (2, 64, 512, 64, 64, 2),    # transposed form on the fused-phase kernel (conv_umma_tmat.cu): one 64-channel chunk, Wj = 256
    (40, 32, 32, 64, 128, 2),   # fused-phase kernel, two chunks, several tiles per CTA and both accumulator sets
    (33, 16, 16, 128, 256, 2),  # fused-phase kernel with 128 output channels: one accumulator set, epilogue sets split the rows
]


@pytest.fixture
def tmat_wide(ops):
    """Run the fused-phase transposed-conv kernel for 128 output channels too (off by default: slower there)."""
    old = ops.set_umma_tmat(3)
    yield
    ops.set_umma_tmat(old)


@pytest.mark.parametrize("case", TMA_CASES)
def test_umma_tma_fprop_and_dgrad(ops, case, tmat_wide):
    B, H, W, Ci, Co, s = case
    x = rnd(B, Ci, H, W, seed=1)
    w = rnd(Co, Ci, 3, 3, seed=2, scale=1.0 / math.sqrt(9 * Ci))
    b = rnd(Co, seed=3, scale=0.1)
    if ops.umma_tma_eligible(B, H, W, Ci, Co, s, 0):
        ref = F.conv2d(x.bfloat16().double(), w.bfloat16().double(), b.double(), stride=s, padding=1)
        stats = torch.empty(2 * Co, dtype=torch.float64, device="cuda")
        y = ops.conv_umma_tma(nhwc(x).cuda().bfloat16(), ops.umma_tma_pack_weight(w.cuda(), s, 0), b.cuda(), stats, Ci, Co, s, 0)
        torch.cuda.synchronize()
        assert rel_err(nchw(y.float().cpu()), ref) < TOL
        assert rel_err(stats.cpu()[:Co], ref.sum(dim=(0, 2, 3))) < 2e-3
        assert rel_err(stats.cpu()[Co:], (ref**2).sum(dim=(0, 2, 3))) < 2e-3
    else:
        assert Ci != 32 and Ci % 64 != 0
    dy = rnd(B, Co, H // s, W // s, seed=4)
    bi = rnd(Ci, seed=5, scale=0.1)
    if ops.umma_tma_eligible(B, H, W, Ci, Co, s, 1):
        ref = F.conv_transpose2d(dy.bfloat16().double(), w.bfloat16().double(), bi.double(), stride=s, padding=1,
                                 output_padding=s - 1)
        stats = torch.empty(2 * Ci, dtype=torch.float64, device="cuda")
        dx = ops.conv_umma_tma(nhwc(dy).cuda().bfloat16(), ops.umma_tma_pack_weight(w.cuda(), s, 1), bi.cuda(), stats, Ci, Co, s, 1)
        torch.cuda.synchronize()
        assert rel_err(nchw(dx.float().cpu()), ref) < TOL
        assert rel_err(stats.cpu()[:Ci], ref.sum(dim=(0, 2, 3))) < 2e-3


# ---------------------------------------------------------------------------
# resident-weight TMA kernel for the narrow layers (conv_umma_rw.cu)
# ---------------------------------------------------------------------------
RW_CASES = [
    # B, H, W (high-res side), Ci, Co, stride, mode
    (2, 32, 32, 32, 32, 2, 0),
    (3, 16, 64, 32, 64, 2, 0),
    (2, 256, 256, 32, 32, 2, 0),   # Wj = 128: one output row per tile
    (5, 16, 16, 32, 64, 2, 0),     # several images per tile, B not a multiple of them
    (1, 16, 8, 32, 32, 1, 0),
    (2, 32, 32, 32, 32, 2, 1),
    (3, 16, 64, 32, 64, 2, 1),     # ConvT 64 -> 32 (128-byte swizzle)
    (5, 16, 16, 32, 32, 2, 1),
    (40, 128, 128, 32, 32, 2, 1),  # persistent loop, TMEM double buffering
    (40, 128, 128, 32, 64, 2, 0),
]


@pytest.mark.parametrize("case", RW_CASES)
def test_umma_rw_conv(ops, case):
    B, H, W, Ci, Co, s, mode = case
    assert ops.umma_rw_eligible(B, H, W, Ci, Co, s, mode)
    w = rnd(Co, Ci, 3, 3, seed=2, scale=1.0 / math.sqrt(9 * (Ci if mode == 0 else Co)))
    wq = w.bfloat16().double()
    if mode == 0:
        x = rnd(B, Ci, H, W, seed=1)
        b = rnd(Co, seed=3, scale=0.1)
        ref = F.conv2d(x.bfloat16().double(), wq, b.double(), stride=s, padding=1)
        inp, n_out = nhwc(x), Co
    else:
        dy = rnd(B, Co, H // s, W // s, seed=4)
        b = rnd(Ci, seed=5, scale=0.1)
        ref = F.conv_transpose2d(dy.bfloat16().double(), wq, b.double(), stride=s, padding=1, output_padding=s - 1)
        inp, n_out = nhwc(dy), Ci
    wpk = ops.umma_rw_pack_weight(w.cuda(), mode)
    stats = torch.empty(2 * n_out, dtype=torch.float64, device="cuda")
    out = ops.conv_umma_rw(inp.cuda().bfloat16(), wpk, b.cuda(), stats, Ci, Co, s, mode)
    torch.cuda.synchronize()
    assert out.shape == (B, ref.shape[2], ref.shape[3], n_out)
    assert rel_err(nchw(out.float().cpu()), ref) < TOL
    assert rel_err(stats.cpu()[:n_out], ref.sum(dim=(0, 2, 3))) < 2e-3
    assert rel_err(stats.cpu()[n_out:], (ref**2).sum(dim=(0, 2, 3))) < 2e-3
    out2 = ops.conv_umma_rw(inp.cuda().bfloat16(), wpk, None, None, Ci, Co, s, mode)  # no bias / stats
    ref2 = ref - b.double().view(1, -1, 1, 1)
    assert rel_err(nchw(out2.float().cpu()), ref2) < TOL


# ---------------------------------------------------------------------------
# fused BatchNorm-backward epilogue (vae_conv_umma_{rw,tma}_bnbwd): out = du, dstats = {sum du, sum du * xhat}
# ---------------------------------------------------------------------------
BNB_CASES = [
    # kind, B, H, W (high-res side of the equivalent conv), Ci, Co, mode
    ("rw", 3, 32, 32, 32, 32, 0),
    ("rw", 2, 16, 64, 32, 64, 0),
    ("rw", 2, 32, 32, 32, 32, 1),
    ("rw", 3, 16, 32, 32, 64, 1),
    ("rw", 16, 128, 128, 32, 32, 0),   # many tiles per CTA, both epilogue sets
    ("tma", 5, 32, 32, 64, 128, 0),
    ("tma", 5, 32, 32, 64, 128, 1),
    ("tma", 6, 16, 16, 128, 256, 0),
    ("tma", 6, 16, 16, 128, 256, 1),
]


@pytest.mark.parametrize("case", BNB_CASES, ids=[f"{c[0]}-m{c[6]}-{c[4]}x{c[5]}-B{c[1]}" for c in BNB_CASES])
def test_conv_with_fused_bn_backward_epilogue(ops, case):
    kind, B, H, W, Ci, Co, mode = case
    s = 2
    w = rnd(Co, Ci, 3, 3, seed=2, scale=1.0 / math.sqrt(9 * (Ci if mode == 0 else Co)))
    if mode == 0:
        x = rnd(B, Ci, H, W, seed=1)
        d_act = F.conv2d(x.bfloat16().double(), w.bfloat16().double(), None, stride=s, padding=1)
        inp, C = nhwc(x), Co
    else:
        x = rnd(B, Co, H // s, W // s, seed=1)
        d_act = F.conv_transpose2d(x.bfloat16().double(), w.bfloat16().double(), None, stride=s, padding=1, output_padding=1)
        inp, C = nhwc(x), Ci
    y = rnd(*d_act.shape, seed=7).bfloat16()
    mean, rstd = rnd(C, seed=8, scale=0.3), rnd(C, seed=9).abs() + 0.5
    gamma, beta = rnd(C, seed=10) + 1.5, rnd(C, seed=11, scale=0.2)
    scale = gamma * rstd
    vec = torch.stack([scale, beta - mean * scale, mean, rstd]).cuda().contiguous()
    cw = ops.ConvWeights(w.cuda(), torch.bfloat16, s)
    wpk = (cw.rw(mode) if kind == "rw" else cw.umma(mode, tma=True))
    du, dstats = ops.conv_umma_bnbwd(kind, inp.cuda().bfloat16(), wpk, nhwc(y).cuda(), vec, beta.cuda(), Ci, Co, s, mode)
    torch.cuda.synchronize()
    y64 = y.double()
    t = y64 - mean.double().view(1, -1, 1, 1)
    u = t * scale.double().view(1, -1, 1, 1) + beta.double().view(1, -1, 1, 1)
    want = torch.where(u > 0, d_act, d_act * 0.01)
    xhat = t * rstd.double().view(1, -1, 1, 1)
    assert rel_err(nchw(du.float().cpu()), want) < TOL
    assert rel_err(dstats.cpu()[:C], want.sum(dim=(0, 2, 3))) < 2e-3
    assert rel_err(dstats.cpu()[C:], (want * xhat).sum(dim=(0, 2, 3))) < 2e-3


# ---------------------------------------------------------------------------
# edge layer (1 input channel) on tensor cores: conv_edge_tc.cu inside vae_conv_fprop
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("case", [(4, 128, 128, 32), (3, 64, 128, 32), (2, 128, 256, 64), (64, 128, 128, 32)],
                         ids=lambda c: "B%d-%dx%d-N%d" % c)
def test_edge_layer_tensor_core_forward(ops, case):
    B, H, W, N = case
    g = torch.Generator().manual_seed(3)
    x = (torch.rand(B, 1, H, W, generator=g) < 0.1).float()          # piano-roll: {0, 1}
    x[0] = torch.rand(1, H, W, generator=g)                            # one grey image: inputs are rounded to bf16 there
    w = rnd(N, 1, 3, 3, seed=2, scale=1.0 / 3.0)
    b = rnd(N, seed=3, scale=0.1)
    cw = ops.ConvWeights(w.cuda(), torch.bfloat16, 2)
    outs = []
    for tc in (True, False):
        old = ops.set_edge_tc(tc)
        try:
            stats = torch.empty(2 * N, dtype=torch.float64, device="cuda")
            y = ops.conv_forward(x.view(B, H, W, 1).cuda(), cw, b.cuda(), stats, Ci=1)
            torch.cuda.synchronize()
        finally:
            ops.set_edge_tc(old)
        outs.append((y.float().cpu(), stats.cpu()))
    ref = F.conv2d(x.bfloat16().double(), w.bfloat16().double(), b.double(), stride=2, padding=1)
    for y, stats in outs:
        assert y.shape == (B, H // 2, W // 2, N)
        assert rel_err(nchw(y), ref) < TOL
        assert rel_err(stats[:N], ref.sum(dim=(0, 2, 3))) < 2e-3
        assert rel_err(stats[N:], (ref**2).sum(dim=(0, 2, 3))) < 2e-3
    # binary images: both kernels see exactly the same operands
    assert rel_err(outs[0][0][1:], outs[1][0][1:]) < 4e-3

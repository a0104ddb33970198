"""Thin torch-tensor wrappers over the C ABI (include/vae_b200.h).

Every wrapper checks device / dtype / contiguity, passes raw device pointers
and the CURRENT torch CUDA stream, and raises RuntimeError on failure.  Device
memory always comes from torch's caching allocator (SURVEY.md §8b "Ownership").
"""

from __future__ import annotations

from typing import Optional

import math

import torch

from . import _lib

F32, BF16 = 0, 1
_CODES = {torch.float32: F32, torch.bfloat16: BF16}
LRELU_SLOPE = 0.01  # nn.LeakyReLU() default, models.py:47
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# ---------------------------------------------------------------- call plumbing
LAUNCHES = 0  # number of C-ABI kernel launches issued by this process
_TRACE = None  # optional Timeline


class Timeline:
    """CUDA-event timing of individual C-ABI launches on the launching stream.

    ``only`` restricts the bracketing to one (name, tag) so the timed region of
    bench.py is perturbed by two event records per step, not hundreds."""

    def __init__(self, only=None):
        self.only = only
        self.records = {}  # (name, tag) -> list[(start, end)]

    def run(self, name, fn, args):
        tag = tuple(a for a in args if isinstance(a, int) and not isinstance(a, bool) and -1 <= a < (1 << 31))
        key = (name, tag)
        if self.only is not None and key != self.only:
            return fn(*args)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = fn(*args)
        e.record()
        self.records.setdefault(key, []).append((s, e))
        return rc

    def summary(self):
        """-> {(name, tag): (n_launches, total_ms)} ; call after a synchronize."""
        return {k: (len(v), sum(s.elapsed_time(e) for s, e in v)) for k, v in self.records.items()}


def set_timeline(tl):
    global _TRACE
    _TRACE = tl


class _Proxy:
    def __init__(self, lib):
        self._lib = lib

    _QUERIES = frozenset({"vae_abi_version", "vae_last_error", "vae_has_umma", "vae_set_umma", "vae_set_umma_wgrad",
                          "vae_umma_packed_elems", "vae_umma_wgrad_supported", "vae_umma_set_trace",
                          "vae_set_umma_tma", "vae_umma_tma_eligible", "vae_umma_tma_packed_elems",
                          "vae_final_block_supported", "vae_set_umma_rw", "vae_umma_rw_eligible",
                          "vae_umma_rw_packed_elems", "vae_set_wgrad_tma", "vae_wgrad_tma_eligible", "vae_heads_supported", "vae_set_final_tc",
                          "vae_final_tc_supported", "vae_linear_tc_supported", "vae_set_edge_tc", "vae_set_umma_tmat", "vae_set_reserved_sms", "vae_set_zero_arena"})

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if name in self._QUERIES:  # host-only queries: not kernel launches
            setattr(self, name, fn)
            return fn

        def call(*args):
            global LAUNCHES
            LAUNCHES += 1
            tr = _TRACE
            return fn(*args) if tr is None else tr.run(name, fn, args)

        setattr(self, name, call)
        return call


_PROXY = None


def _L():
    global _PROXY
    if _PROXY is None:
        _PROXY = _Proxy(_lib.load())
    return _PROXY


def code(t) -> int:
    dt = t if isinstance(t, torch.dtype) else t.dtype
    try:
        return _CODES[dt]
    except KeyError:
        raise RuntimeError(f"torch_vae_b200: unsupported dtype {dt} (float32 / bfloat16 only)")


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("torch_vae_b200 kernels need CUDA tensors (there is no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("torch_vae_b200 kernels need contiguous tensors")
    if t.device.index != torch.cuda.current_device():
        # kernels are enqueued on the CURRENT device's current stream (one process per GPU, torchrun model):
        # a tensor of another device would be dereferenced by the wrong GPU
        raise RuntimeError(f"torch_vae_b200: tensor lives on {t.device} but the current CUDA device is "
                           f"cuda:{torch.cuda.current_device()}; call torch.cuda.set_device() first")
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32(t: Optional[torch.Tensor], what: str):
    if t is not None and t.dtype != torch.float32:
        raise RuntimeError(f"{what} must be float32, got {t.dtype}")
    return _p(t)


# ------------------------------------------------------------------ conv engine
def pack_conv_weight(w: torch.Tensor, dtype: torch.dtype, want_wp=True, want_wt=True):
    """OIHW fp32 -> (wp[Co,3,3,Ci], wt[Ci,3,3,Co]) in ``dtype``."""
    Co, Ci = w.shape[0], w.shape[1]
    wp = torch.empty((Co, 3, 3, Ci), dtype=dtype, device=w.device) if want_wp else None
    wt = torch.empty((Ci, 3, 3, Co), dtype=dtype, device=w.device) if want_wt else None
    rc = _L().vae_pack_conv_weight(_f32(w, "conv weight"), _p(wp), _p(wt), Co, Ci, code(dtype), _stream())
    _lib.check(rc, "vae_pack_conv_weight")
    return wp, wt


def conv_fprop(x, wp, bias, stride: int, out_dtype, stats=None, Ci: Optional[int] = None):
    """x NHWC [B,Hi,Wi,Ci] (or [B,1,H,W]/[B,H,W] when Ci == 1) -> y NHWC."""
    if Ci is None:
        Ci = x.shape[-1]
    if Ci == 1 and x.dim() == 4 and x.shape[1] == 1 and x.shape[-1] != 1:
        B, Hi, Wi = x.shape[0], x.shape[2], x.shape[3]  # NCHW with C == 1 is NHWC
    else:
        B, Hi, Wi = x.shape[0], x.shape[1], x.shape[2]
    Co = wp.shape[0]
    y = torch.empty((B, Hi // stride, Wi // stride, Co), dtype=out_dtype, device=x.device)
    rc = _L().vae_conv_fprop(
        _p(x), _p(wp), _f32(bias, "bias"), _p(y), _p(stats), B, Hi, Wi, Ci, Co, stride,
        code(x), code(out_dtype), _stream(),
    )
    _lib.check(rc, "vae_conv_fprop")
    return y


def conv_dgrad(dy, wt, bias, stride: int, stats=None):
    """dy NHWC [B,Ho,Wo,Co], wt [Ci,3,3,Co] -> dx NHWC [B,Ho*s,Wo*s,Ci]."""
    B, Ho, Wo, Co = dy.shape
    Ci = wt.shape[0]
    dx = torch.empty((B, Ho * stride, Wo * stride, Ci), dtype=dy.dtype, device=dy.device)
    rc = _L().vae_conv_dgrad(
        _p(dy), _p(wt), _f32(bias, "bias"), _p(dx), _p(stats), B, Ho * stride, Wo * stride, Ci, Co,
        stride, code(dy), _stream(),
    )
    _lib.check(rc, "vae_conv_dgrad")
    return dx


def conv_wgrad(x, dy, stride: int, Ci: int, Co: int, Hi: int, Wi: int, out=None):
    """-> dw [Co,Ci,3,3] fp32 (reference OIHW layout)."""
    B = dy.shape[0]
    dw = out if out is not None else torch.empty((Co, Ci, 3, 3), dtype=torch.float32, device=dy.device)
    rc = _L().vae_conv_wgrad(_p(x), _p(dy), _p(dw), B, Hi, Wi, Ci, Co, stride, code(x), code(dy), _stream())
    _lib.check(rc, "vae_conv_wgrad")
    return dw


# ------------------------------------------------------------------- batch norm
def bn_finalize(stats, gamma, beta, running_mean, running_var, count: int, training: bool,
                momentum: float = BN_MOMENTUM, eps: float = BN_EPS):
    C = gamma.numel()
    vec = torch.empty((4, C), dtype=torch.float32, device=gamma.device)  # scale, shift, mean, rstd
    rc = _L().vae_bn_finalize(
        _p(stats), _f32(gamma, "bn weight"), _f32(beta, "bn bias"), _f32(running_mean, "running_mean"),
        _f32(running_var, "running_var"), vec[0].data_ptr(), vec[1].data_ptr(), vec[2].data_ptr(),
        vec[3].data_ptr(), C, count, momentum, eps, 1 if training else 0, _stream(),
    )
    _lib.check(rc, "vae_bn_finalize")
    return vec


def bn_finalize_apply_lrelu(y, stats, gamma, beta, running_mean, running_var, num_batches_tracked, count: int,
                            training: bool, momentum: float = BN_MOMENTUM, eps: float = BN_EPS,
                            slope: float = LRELU_SLOPE):
    """BatchNorm statistics -> vectors AND normalise + LeakyReLU in one launch -> (vec[4,C], act)."""
    C = y.shape[-1]
    vec = torch.empty((4, C), dtype=torch.float32, device=y.device)
    act = torch.empty_like(y)
    rc = _L().vae_bn_finalize_apply_lrelu(
        _p(y), _p(stats), _f32(gamma, "bn weight"), _f32(beta, "bn bias"), _f32(running_mean, "running_mean"),
        _f32(running_var, "running_var"), _p(num_batches_tracked), _p(vec), _p(act), y.numel() // C, C, count,
        momentum, eps, 1 if training else 0, slope, code(y), _stream(),
    )
    _lib.check(rc, "vae_bn_finalize_apply_lrelu")
    return vec, act


def bn_apply_lrelu(y, vec, beta, slope: float = LRELU_SLOPE):
    C = y.shape[-1]
    act = torch.empty_like(y)
    rc = _L().vae_bn_apply_lrelu(
        _p(y), vec[0].data_ptr(), vec[2].data_ptr(), _f32(beta, "bn bias"), _p(act), y.numel() // C, C, slope,
        code(y), _stream()
    )
    _lib.check(rc, "vae_bn_apply_lrelu")
    return act


# One cooperative launch for both passes of the BatchNorm backward when the tensors stay in the L2 (bn_lrelu_backward).
# Measured SLOWER on the B200 (128 channels @16x16, B = 512: 69 us against 25 + 23 us for the two kernels): the
# cooperative launch and the grid-wide arrival cost more than the second HBM read they save.  Opt-in; DESIGN.md 6.1.
FUSE_BN_BWD_PASSES = False


def _small_outs(outs, C, device):
    """(dgamma, dbeta, dbias) destinations: caller-provided fp32 [C] tensors (gradient-bucket views) or fresh ones."""
    if outs is not None:
        for t in outs:
            if t.dtype != torch.float32 or t.numel() != C:
                raise RuntimeError("gradient destination must be float32 with one element per channel")
        return outs
    small = torch.empty((3, C), dtype=torch.float32, device=device)
    return small[0], small[1], small[2]


def bn_lrelu_backward(dact, y, vec, gamma, beta, training: bool, slope: float = LRELU_SLOPE, outs=None, dstats=None):
    """-> (dy, dgamma, dbeta, dbias); ``outs`` = optional (dgamma, dbeta, dbias) destinations.  With ``dstats``
    (from a conv_*(bnbwd=...) call) ``dact`` already holds du and the reduction pass is skipped."""
    C = y.shape[-1]
    n_pix = y.numel() // C
    lib = _L()
    is_du = dstats is not None
    if not is_du and FUSE_BN_BWD_PASSES and 2 * y.numel() * y.element_size() <= 96 << 20:
        # small enough to stay in the L2 between the passes: one cooperative launch (reduce, grid arrival, apply)
        dstats = torch.empty(2 * C + 1, dtype=torch.float64, device=y.device)
        dy = torch.empty_like(y)
        dgamma, dbeta, dbias = _small_outs(outs, C, y.device)
        rc = lib.vae_bn_lrelu_bwd_fused(
            _p(dact), _p(y), vec[0].data_ptr(), _f32(beta, "bn bias"), vec[2].data_ptr(), vec[3].data_ptr(),
            _f32(gamma, "bn weight"), _p(dstats), _p(dy), _p(dgamma), _p(dbeta), _p(dbias), n_pix, C, slope,
            1 if training else 0, code(y), _stream(),
        )
        if rc == 0:
            return dy, dgamma, dbeta, dbias
        if rc != -3:   # -3: VAE_ERR_UNSUPPORTED, take the two-pass path
            _lib.check(rc, "vae_bn_lrelu_bwd_fused")
        dstats = None
    if not is_du:
        dstats = scratch(2 * C, torch.float64, y.device)
        rc = lib.vae_bn_lrelu_bwd_reduce(
            _p(dact), _p(y), vec[0].data_ptr(), _f32(beta, "bn bias"), vec[2].data_ptr(), vec[3].data_ptr(),
            _p(dstats), n_pix, C, slope, code(y), _stream(),
        )
        _lib.check(rc, "vae_bn_lrelu_bwd_reduce")
    dy = torch.empty_like(y)
    dgamma, dbeta, dbias = _small_outs(outs, C, y.device)
    rc = lib.vae_bn_lrelu_bwd_apply(
        _p(dact), _p(y), vec[0].data_ptr(), _f32(beta, "bn bias"), vec[2].data_ptr(), vec[3].data_ptr(),
        _f32(gamma, "bn weight"), _p(dstats), _p(dy), _p(dgamma), _p(dbeta), _p(dbias), n_pix, C, slope,
        (1 if training else 0) | (2 if is_du else 0), code(y), _stream(),
    )
    _lib.check(rc, "vae_bn_lrelu_bwd_apply")
    return dy, dgamma, dbeta, dbias


# ------------------------------------------------- fused output block (models.py:76-82)
def final_block_supported(B: int, H: int, W: int, C: int) -> bool:
    return bool(_L().vae_final_block_supported(B, H, W, C))


def final_tc_supported(y) -> bool:
    B, H, W, C = y.shape
    return y.dtype == torch.bfloat16 and umma_available() and bool(_L().vae_final_tc_supported(B, H, W, C, code(y)))


def set_final_tc(enable: bool) -> bool:
    return bool(_L().vae_set_final_tc(1 if enable else 0))


def final_conv_fwd(y, vec, beta, w_last, b_last, slope: float = LRELU_SLOPE):
    """y [B,H,W,C] raw ConvT output -> (logits [B,1,H,W] in y.dtype, probs [B,1,H,W] fp32):
    BatchNorm apply + LeakyReLU + Conv2d(C->1) + Sigmoid in one pass over y."""
    B, H, W, C = y.shape
    logits = torch.empty((B, 1, H, W), dtype=y.dtype, device=y.device)
    probs = torch.empty((B, 1, H, W), dtype=torch.float32, device=y.device)
    if final_tc_supported(y):
        rc = _L().vae_final_conv_fwd_tc(
            _p(y), vec[2].data_ptr(), vec[0].data_ptr(), _f32(beta, "bn bias"), _f32(w_last, "final conv weight"),
            _f32(b_last, "final conv bias"), _p(logits), _p(probs), B, H, W, C, slope, _stream(),
        )
        _lib.check(rc, "vae_final_conv_fwd_tc")
        return logits, probs
    rc = _L().vae_final_conv_fwd(
        _p(y), vec[2].data_ptr(), vec[0].data_ptr(), _f32(beta, "bn bias"), _f32(w_last, "final conv weight"),
        _f32(b_last, "final conv bias"), _p(logits), _p(probs), B, H, W, C, slope, code(y), _stream(),
    )
    _lib.check(rc, "vae_final_conv_fwd")
    return logits, probs


def final_block_backward(dlogits, y, vec, gamma, beta, w_last, training: bool, slope: float = LRELU_SLOPE,
                         outs=None, last_outs=None):
    """-> (dy, dgamma, dbeta, dbias, dw_last, db_last): backward of Conv2d(C->1) + LeakyReLU +
    BatchNorm in two passes over y (reduce, apply).  ``outs`` = optional (dgamma, dbeta, dbias)
    destinations, ``last_outs`` = optional (dw_last [1,C,3,3], db_last [1]) destinations."""
    B, H, W, C = y.shape
    lib = _L()
    dstats = scratch(2 * C + 1, torch.float64, y.device)
    dw_last = last_outs[0] if last_outs is not None else torch.empty((1, C, 3, 3), dtype=torch.float32, device=y.device)
    mean, scale, rstd = vec[2].data_ptr(), vec[0].data_ptr(), vec[3].data_ptr()
    tc = final_tc_supported(y)
    if tc:
        rc = lib.vae_final_bwd_reduce_tc(
            _p(dlogits), _p(y), mean, scale, _f32(beta, "bn bias"), rstd, _f32(w_last, "final conv weight"),
            _p(dstats), _p(dw_last), B, H, W, C, slope, _stream(),
        )
    else:
        rc = lib.vae_final_bwd_reduce(
            _p(dlogits), _p(y), mean, scale, _f32(beta, "bn bias"), rstd, _f32(w_last, "final conv weight"),
            _p(dstats), _p(dw_last), B, H, W, C, slope, code(y), _stream(),
        )
    _lib.check(rc, "vae_final_bwd_reduce")
    dy = torch.empty_like(y)
    dgamma, dbeta, dbias = _small_outs(outs, C, y.device)
    db_last = last_outs[1] if last_outs is not None else torch.empty(1, dtype=torch.float32, device=y.device)
    if tc:
        rc = lib.vae_final_bwd_apply_tc(
            _p(dlogits), _p(y), mean, scale, _f32(beta, "bn bias"), rstd, _f32(gamma, "bn weight"),
            _f32(w_last, "final conv weight"), _p(dstats), _p(dy), _p(dgamma), _p(dbeta), _p(dbias),
            _p(db_last), B, H, W, C, slope, 1 if training else 0, _stream(),
        )
    else:
        rc = lib.vae_final_bwd_apply(
            _p(dlogits), _p(y), mean, scale, _f32(beta, "bn bias"), rstd, _f32(gamma, "bn weight"),
            _f32(w_last, "final conv weight"), _p(dstats), _p(dy), _p(dgamma), _p(dbeta), _p(dbias),
            _p(db_last), B, H, W, C, slope, 1 if training else 0, code(y), _stream(),
        )
    _lib.check(rc, "vae_final_bwd_apply")
    return dy, dgamma, dbeta, dbias, dw_last, db_last


# ---------------------------------------------------------------- layout glue
def nhwc_to_nchw(src, out_dtype=None):
    """[B,H,W,C] -> [B, C*H*W] (the reference's C-major flatten, models.py:133)."""
    B, H, W, Cc = src.shape
    out_dtype = out_dtype or src.dtype
    dst = torch.empty((B, Cc * H * W), dtype=out_dtype, device=src.device)
    rc = _L().vae_nhwc_to_nchw(_p(src), _p(dst), B, H * W, Cc, code(src), code(out_dtype), _stream())
    _lib.check(rc, "vae_nhwc_to_nchw")
    return dst


def nchw_to_nhwc(src, C: int, H: int, W: int, out_dtype=None):
    """[B, C*H*W] -> [B,H,W,C] (the reference's view, models.py:166)."""
    B = src.shape[0]
    out_dtype = out_dtype or src.dtype
    dst = torch.empty((B, H, W, C), dtype=out_dtype, device=src.device)
    rc = _L().vae_nchw_to_nhwc(_p(src), _p(dst), B, H * W, C, code(src), code(out_dtype), _stream())
    _lib.check(rc, "vae_nchw_to_nhwc")
    return dst


# ---------------------------------------------------------------------- linear
def gemm(A, Bm, out, M, N, K, sa, sb, sc, bias=None, accumulate=False):
    """out[M,N] (+)= A[M,K] @ B[K,N] with explicit (row, col) element strides."""
    rc = _L().vae_gemm(
        _p(A), _p(Bm), _p(out), _f32(bias, "bias"), M, N, K, sa[0], sa[1], sb[0], sb[1], sc[0], sc[1],
        code(A), code(Bm), code(out), 1 if accumulate else 0, _stream(),
    )
    _lib.check(rc, "vae_gemm")
    return out


def heads_supported(F: int, L: int) -> bool:
    return bool(_L().vae_heads_supported(F, L))


def skinny_reduce(A, W, bias, wsn: int, wsk: int, N: int):
    """out[M,N] fp32 = A[M,K] . W(n,k) + bias, W(n,k) = W.flat[n*wsn + k*wsk]."""
    M, K = A.shape
    out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    rc = _L().vae_skinny_reduce(_p(A), _f32(W, "weight"), _f32(bias, "bias"), _p(out), M, N, K, wsn, wsk, code(A),
                                _stream())
    _lib.check(rc, "vae_skinny_reduce")
    return out


def skinny_reduce_pair(A, W0, bias0, W1, bias1, wsn: int, wsk: int, N: int):
    """(A . W0 + bias0, A . W1 + bias1) in ONE pass over A: fc_mu and fc_var share their input (models.py:137,141)."""
    M, K = A.shape
    out = torch.empty((2, M, N), dtype=torch.float32, device=A.device)
    rc = _L().vae_skinny_reduce_pair(_p(A), _f32(W0, "weight"), _f32(bias0, "bias"), out[0].data_ptr(), _f32(W1, "weight"),
                                     _f32(bias1, "bias"), out[1].data_ptr(), M, N, K, wsn, wsk, code(A), _stream())
    _lib.check(rc, "vae_skinny_reduce_pair")
    return out[0], out[1]


def skinny_wgrad(D, A, g_shape, gsn: int, gsk: int, want_colsum: bool, out=None, colsum_out=None):
    """G(n,k) = sum_m D[m,n] A[m,k] laid out as G.flat[n*gsn + k*gsk]; colsum[k] = sum_m A[m,k].
    ``out`` / ``colsum_out``: optional destinations (gradient-bucket views); they are zeroed here."""
    M, K = A.shape
    N = D.shape[1]
    if out is not None:
        G = out.view(g_shape)
        G.zero_()
        cs = None
        if want_colsum:
            cs = colsum_out if colsum_out is not None else torch.empty(K, dtype=torch.float32, device=A.device)
            cs.zero_()
    else:
        buf = torch.zeros(N * K + (K if want_colsum else 0), dtype=torch.float32, device=A.device)
        G = buf[: N * K].view(g_shape)
        cs = buf[N * K:] if want_colsum else None
    rc = _L().vae_skinny_wgrad(_f32(D, "latent gradient"), _p(A), G.data_ptr(), cs.data_ptr() if want_colsum else None,
                               M, N, K, gsn, gsk, code(A), _stream())
    _lib.check(rc, "vae_skinny_wgrad")
    return G, cs


def skinny_expand(D0, W0, D1, W1, bias, wsn: int, wsk: int, K: int, out_dtype):
    """out[M,K] = bias + D0 . W0 (+ D1 . W1)."""
    M, N = D0.shape
    out = torch.empty((M, K), dtype=out_dtype, device=D0.device)
    rc = _L().vae_skinny_expand(_f32(D0, "latent"), _f32(W0, "weight"), _f32(D1, "latent"), _f32(W1, "weight"),
                                _f32(bias, "bias"), _p(out), M, N, K, wsn, wsk, code(out_dtype), _stream())
    _lib.check(rc, "vae_skinny_expand")
    return out


def linear_tc_supported(M: int, L: int, F: int) -> bool:
    """Tensor-core Linear kernels (linear_tc.cu): batch <= 64, latent 32..256, F features (bf16 activations)."""
    return bool(_L().vae_linear_tc_supported(M, L, F))


def linear_fwd(x, w, b, out_dtype):
    """x[B,K] @ w[N,K]^T + b -> [B,N]   (aten::addmm, models.py:137,141,162)."""
    Bn, K = x.shape
    N = w.shape[0]
    out = torch.empty((Bn, N), dtype=out_dtype, device=x.device)
    if x.dtype == torch.bfloat16 and out_dtype == torch.float32 and K > N and linear_tc_supported(Bn, N, K):
        # fc_mu / fc_var: reduction over the features, W[L][F]
        rc = _L().vae_linear_reduce_tc(_p(x), _f32(w, "weight"), _f32(b, "bias"), _p(out), Bn, N, K, 0, _stream())
        _lib.check(rc, "vae_linear_reduce_tc")
        return out
    if x.dtype == torch.float32 and out_dtype == torch.bfloat16 and N > K and linear_tc_supported(Bn, K, N):
        # decoder_input: expansion of the latent, W[F][L]
        rc = _L().vae_linear_expand_tc(_p(x), _f32(w, "weight"), None, None, _f32(b, "bias"), _p(out), Bn, K, N, 1,
                                       code(out_dtype), _stream())
        _lib.check(rc, "vae_linear_expand_tc")
        return out
    return gemm(x, w, out, Bn, N, K, (K, 1), (1, K), (N, 1), bias=b)


def linear_bwd_input(dout, w, out=None, out_dtype=torch.float32, accumulate=False):
    """dout[B,N] @ w[N,K] -> dx[B,K]."""
    Bn, N = dout.shape
    K = w.shape[1]
    if out is None:
        out = torch.empty((Bn, K), dtype=out_dtype, device=dout.device)
    if not accumulate and dout.dtype == torch.bfloat16 and out.dtype == torch.float32 and N > K and linear_tc_supported(Bn, K, N):
        # latent gradient of decoder_input: reduction over the features, W[F][L]
        rc = _L().vae_linear_reduce_tc(_p(dout), _f32(w, "weight"), None, _p(out), Bn, K, N, 1, _stream())
        _lib.check(rc, "vae_linear_reduce_tc")
        return out
    return gemm(dout, w, out, Bn, K, N, (N, 1), (K, 1), (K, 1), accumulate=accumulate)


def linear_bwd_input_pair(d0, w0, d1, w1, out_dtype):
    """d0[B,L] @ w0[L,F] + d1[B,L] @ w1[L,F] -> [B,F]: the input gradient of two heads that share their input
    (fc_mu + fc_var, models.py:137,141), one pass when the tensor-core kernel serves the shape."""
    Bn, L = d0.shape
    F_ = w0.shape[1]
    if d0.dtype == torch.float32 and d1.dtype == torch.float32 and out_dtype == torch.bfloat16 and linear_tc_supported(Bn, L, F_):
        out = torch.empty((Bn, F_), dtype=out_dtype, device=d0.device)
        rc = _L().vae_linear_expand_tc(_p(d0), _f32(w0, "weight"), _p(d1), _f32(w1, "weight"), None, _p(out), Bn, L, F_, 0,
                                       code(out_dtype), _stream())
        _lib.check(rc, "vae_linear_expand_tc")
        return out
    out = linear_bwd_input(d0, w0, out_dtype=out_dtype)
    return linear_bwd_input(d1, w1, out=out, accumulate=True)


def linear_bwd_weight(dout, x, out=None):
    """dout[B,N]^T @ x[B,K] -> dw[N,K] fp32."""
    Bn, N = dout.shape
    K = x.shape[1]
    dw = out if out is not None else torch.empty((N, K), dtype=torch.float32, device=x.device)
    if dout.dtype == torch.float32 and x.dtype == torch.bfloat16 and K > N and linear_tc_supported(Bn, N, K):
        # fc_mu / fc_var: G[L][F] = dmu^T . pre
        rc = _L().vae_linear_wgrad_tc(_p(dout), _p(x), _p(dw), Bn, N, K, 0, _stream())
        _lib.check(rc, "vae_linear_wgrad_tc")
        return dw
    if dout.dtype == torch.bfloat16 and x.dtype == torch.float32 and N > K and linear_tc_supported(Bn, K, N):
        # decoder_input: G[F][L] = d_d^T . z
        rc = _L().vae_linear_wgrad_tc(_p(x), _p(dout), _p(dw), Bn, K, N, 1, _stream())
        _lib.check(rc, "vae_linear_wgrad_tc")
        return dw
    return gemm(dout, x, dw, N, K, Bn, (1, N), (K, 1), (K, 1))


def colsum(A, out=None):
    M, N = A.shape
    if out is None:
        out = torch.empty(N, dtype=torch.float32, device=A.device)
    rc = _L().vae_colsum(_p(A), _p(out), M, N, code(A), _stream())
    _lib.check(rc, "vae_colsum")
    return out


# --------------------------------------------------------------- reparam / rng
def philox_randn(n: int, seed: int, offset: int, device) -> torch.Tensor:
    out = torch.empty(n, dtype=torch.float32, device=device)
    rc = _L().vae_philox_randn(_p(out), n, seed, offset, _stream())
    _lib.check(rc, "vae_philox_randn")
    return out


def reparam_fwd(mu, logvar, eps, seed: int = 0, offset: int = 0, dev_offset=None):
    z = torch.empty_like(mu)
    eps_out = None if eps is not None else torch.empty_like(mu)
    rc = _L().vae_reparam_fwd(
        _f32(mu, "mu"), _f32(logvar, "log_var"), _f32(eps, "eps"), _p(eps_out), _p(z), mu.numel(),
        seed, offset, _p(dev_offset), _stream(),
    )
    _lib.check(rc, "vae_reparam_fwd")
    return z, (eps if eps is not None else eps_out)


def reparam_bwd(dz, eps, logvar):
    dmu = torch.empty_like(dz)
    dlv = torch.empty_like(dz)
    rc = _L().vae_reparam_bwd(_f32(dz, "dz"), _p(eps), _p(logvar), _p(dmu), _p(dlv), dz.numel(), 0, _stream())
    _lib.check(rc, "vae_reparam_bwd")
    return dmu, dlv


# ------------------------------------------------------------------------ ELBO
def elbo_fwd(logits, target, mu, logvar, kld_weight: float, input_is_prob: bool):
    losses = torch.empty(3, dtype=torch.float32, device=logits.device)
    scratch_ = scratch(4, torch.float64, logits.device)
    B, L = mu.shape
    rc = _L().vae_elbo_fwd(
        _p(logits), _f32(target, "input"), _f32(mu, "mu"), _f32(logvar, "log_var"), _p(losses),
        _p(scratch_), logits.numel(), B, L, kld_weight, code(logits), 1 if input_is_prob else 0, _stream(),
    )
    _lib.check(rc, "vae_elbo_fwd")
    return losses


def elbo_bwd(logits, target, mu, logvar, gscale, kld_weight: float, input_is_prob: bool):
    dlogits = torch.empty_like(logits)
    dmu = torch.empty_like(mu)
    dlv = torch.empty_like(mu)
    B, L = mu.shape
    rc = _L().vae_elbo_bwd(
        _p(logits), _f32(target, "input"), _p(mu), _p(logvar), _f32(gscale, "grad"), _p(dlogits),
        _p(dmu), _p(dlv), logits.numel(), B, L, kld_weight, code(logits), 1 if input_is_prob else 0,
        _stream(),
    )
    _lib.check(rc, "vae_elbo_bwd")
    return dlogits, dmu, dlv


def elbo_fused(logits, target, mu, logvar, kld_weight: float, input_is_prob: bool):
    """Losses and, for an upstream gradient of 1, (dlogits, dmu, dlogvar) in one pass."""
    losses = torch.empty(3, dtype=torch.float32, device=logits.device)
    scratch_ = scratch(4, torch.float64, logits.device)
    dlogits = torch.empty_like(logits)
    dmu = torch.empty_like(mu)
    dlv = torch.empty_like(mu)
    B, L = mu.shape
    rc = _L().vae_elbo_fused(
        _p(logits), _f32(target, "input"), _f32(mu, "mu"), _f32(logvar, "log_var"), _p(losses), _p(scratch_),
        _p(dlogits), _p(dmu), _p(dlv), logits.numel(), B, L, kld_weight, code(logits),
        1 if input_is_prob else 0, _stream(),
    )
    _lib.check(rc, "vae_elbo_fused")
    return losses, dlogits, dmu, dlv


def elbo_scale_grads(gscale, dlogits, dmu, dlv):
    """In-place ``*= gscale`` (device scalar) of the cached ELBO gradients; free when it is 1."""
    rc = _L().vae_elbo_scale_grads(
        _f32(gscale, "grad"), _p(dlogits), _p(dmu), _p(dlv), dlogits.numel(), dmu.numel(), code(dlogits), _stream(),
    )
    _lib.check(rc, "vae_elbo_scale_grads")


def sigmoid(logits):
    probs = torch.empty(logits.shape, dtype=torch.float32, device=logits.device)
    rc = _L().vae_sigmoid(_p(logits), _p(probs), logits.numel(), code(logits), _stream())
    _lib.check(rc, "vae_sigmoid")
    return probs


def sigmoid_bwd(dprobs, probs, dlogits=None, dtype=torch.float32):
    acc = dlogits is not None
    if dlogits is None:
        dlogits = torch.empty(probs.shape, dtype=dtype, device=probs.device)
    rc = _L().vae_sigmoid_bwd(
        _f32(dprobs, "grad of output"), _f32(probs, "output"), _p(dlogits), probs.numel(),
        code(dlogits), 1 if acc else 0, _stream(),
    )
    _lib.check(rc, "vae_sigmoid_bwd")
    return dlogits


# ------------------------------------------------------------------ transforms
def roll_transform(rolls, H: int, W: int, crop_mode: int, max_pitch_shift: int,
                   binarize_threshold: float, mean: float, std: float, seed: int, offset: int, out=None):
    if rolls.dtype != torch.uint8 or rolls.dim() != 3:
        raise RuntimeError("roll_transform expects uint8 rolls of shape [B, Hs, Ws]")
    B, Hs, Ws = rolls.shape
    if out is None:
        out = torch.empty((B, 1, H, W), dtype=torch.float32, device=rolls.device)
    elif out.shape != (B, 1, H, W) or out.dtype != torch.float32 or not out.is_contiguous():
        raise RuntimeError("roll_transform: out must be a contiguous float32 [B, 1, H, W] tensor")
    rc = _L().vae_roll_transform(
        _p(rolls), _p(out), B, Hs, Ws, H, W, crop_mode, max_pitch_shift, binarize_threshold, mean, std,
        seed, offset, _stream(),
    )
    _lib.check(rc, "vae_roll_transform")
    return out


# ------------------------------------------------------- tcgen05 / TMEM conv path
def umma_available() -> bool:
    return bool(_L().vae_has_umma())


def set_umma(enable: bool) -> bool:
    """Enable/disable the tcgen05 path (tests cross-check it against the CUDA-core
    kernels).  Returns the previous setting."""
    return bool(_L().vae_set_umma(1 if enable else 0))


def umma_packed_elems(Ci: int, Co: int, stride: int, mode: int) -> int:
    return int(_L().vae_umma_packed_elems(Ci, Co, stride, mode))


def umma_pack_weight(w: torch.Tensor, stride: int, mode: int) -> Optional[torch.Tensor]:
    """fp32 OIHW weight -> pre-swizzled bf16 smem images for vae_conv_umma
    (mode 0: forward conv, mode 1: data gradient / ConvTranspose forward).
    Returns None when the shape is not served by the tcgen05 path."""
    Co, Ci = w.shape[0], w.shape[1]
    n = umma_packed_elems(Ci, Co, stride, mode)
    if n == 0:
        return None
    out = torch.empty(n, dtype=torch.bfloat16, device=w.device)
    rc = _L().vae_umma_pack_weight(_f32(w, "conv weight"), _p(out), Ci, Co, stride, mode, _stream())
    _lib.check(rc, "vae_umma_pack_weight")
    return out


def conv_umma(inp, wpk, bias, stats, Ci: int, Co: int, stride: int, mode: int, pro=None,
              slope: float = LRELU_SLOPE):
    """mode 0: inp = x [B,Hi,Wi,Ci] -> y [B,Hi/s,Wi/s,Co];
    mode 1: inp = dy [B,Hi/s,Wi/s,Co] -> dx [B,Hi,Wi,Ci].
    pro = (vec, beta) of the BatchNorm that follows the layer which produced ``inp``:
    its BN+LeakyReLU is applied while the operand is gathered."""
    if inp.dtype != torch.bfloat16:
        raise RuntimeError("conv_umma: bf16 activations only")
    B = inp.shape[0]
    if mode == 0:
        Hi, Wi = inp.shape[1], inp.shape[2]
        out = torch.empty((B, Hi // stride, Wi // stride, Co), dtype=torch.bfloat16, device=inp.device)
    else:
        Hi, Wi = inp.shape[1] * stride, inp.shape[2] * stride
        out = torch.empty((B, Hi, Wi, Ci), dtype=torch.bfloat16, device=inp.device)
    ps = pro[0][0].data_ptr() if pro is not None else None
    pm = pro[0][2].data_ptr() if pro is not None else None
    pb = _f32(pro[1], "bn bias") if pro is not None else None
    rc = _L().vae_conv_umma(_p(inp), _p(wpk), _f32(bias, "bias"), _p(out), _p(stats), ps, pm, pb, slope,
                            B, Hi, Wi, Ci, Co, stride, mode, _stream())
    _lib.check(rc, "vae_conv_umma")
    return out


def umma_tma_eligible(B: int, Hi: int, Wi: int, Ci: int, Co: int, stride: int, mode: int) -> bool:
    return bool(_L().vae_umma_tma_eligible(B, Hi, Wi, Ci, Co, stride, mode))


def umma_tma_pack_weight(w: torch.Tensor, stride: int, mode: int, out: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    Co, Ci = w.shape[0], w.shape[1]
    if out is None:
        n = int(_L().vae_umma_tma_packed_elems(Ci, Co, stride, mode))
        if n == 0:
            return None
        out = torch.empty(n, dtype=torch.bfloat16, device=w.device)
    rc = _L().vae_umma_tma_pack_weight(_f32(w, "conv weight"), _p(out), Ci, Co, stride, mode, _stream())
    _lib.check(rc, "vae_umma_tma_pack_weight")
    return out


def conv_umma_tma(inp, wpk, bias, stats, Ci: int, Co: int, stride: int, mode: int):
    """TMA-fed tcgen05 conv: same contract as conv_umma without the fused prologue."""
    if inp.dtype != torch.bfloat16:
        raise RuntimeError("conv_umma_tma: bf16 activations only")
    B = inp.shape[0]
    if mode == 0:
        Hi, Wi = inp.shape[1], inp.shape[2]
        out = torch.empty((B, Hi // stride, Wi // stride, Co), dtype=torch.bfloat16, device=inp.device)
    else:
        Hi, Wi = inp.shape[1] * stride, inp.shape[2] * stride
        out = torch.empty((B, Hi, Wi, Ci), dtype=torch.bfloat16, device=inp.device)
    rc = _L().vae_conv_umma_tma(_p(inp), _p(wpk), _f32(bias, "bias"), _p(out), _p(stats), B, Hi, Wi, Ci, Co,
                                stride, mode, _stream())
    _lib.check(rc, "vae_conv_umma_tma")
    return out


def umma_rw_eligible(B: int, Hi: int, Wi: int, Ci: int, Co: int, stride: int, mode: int) -> bool:
    return bool(_L().vae_umma_rw_eligible(B, Hi, Wi, Ci, Co, stride, mode))


def set_reserved_sms(n: int) -> int:
    """Leave ``n`` SMs to concurrent NCCL kernels (persistent kernels size their grids for the rest); returns the old value."""
    return int(_L().vae_set_reserved_sms(int(n)))


def set_umma_tmat(flags: int) -> int:
    """Fused-phase transposed conv: bit 0 = 64 output channels (default on), bit 1 = also 128 (slower; tests).  Also
    selects the packed weight layout: flip before packing.  Returns the previous flags."""
    return int(_L().vae_set_umma_tmat(int(flags)))


def set_edge_tc(enable: bool) -> bool:
    return bool(_L().vae_set_edge_tc(1 if enable else 0))


def set_umma_rw(enable: bool) -> bool:
    return bool(_L().vae_set_umma_rw(1 if enable else 0))


def umma_rw_pack_weight(w: torch.Tensor, mode: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    Co, Ci = w.shape[0], w.shape[1]
    if out is None:
        out = torch.empty(9 * Ci * Co, dtype=torch.bfloat16, device=w.device)
    rc = _L().vae_umma_rw_pack_weight(_f32(w, "conv weight"), _p(out), Ci, Co, mode, _stream())
    _lib.check(rc, "vae_umma_rw_pack_weight")
    return out


def conv_umma_rw(inp, wpk, bias, stats, Ci: int, Co: int, stride: int, mode: int):
    """Resident-weight TMA/tcgen05 conv for the narrow layers (same contract as conv_umma_tma);
    mode 1 computes the four output-parity phases of the transposed conv per tile."""
    if inp.dtype != torch.bfloat16:
        raise RuntimeError("conv_umma_rw: bf16 activations only")
    B = inp.shape[0]
    if mode == 0:
        Hi, Wi = inp.shape[1], inp.shape[2]
        out = torch.empty((B, Hi // stride, Wi // stride, Co), dtype=torch.bfloat16, device=inp.device)
    else:
        Hi, Wi = inp.shape[1] * stride, inp.shape[2] * stride
        out = torch.empty((B, Hi, Wi, Ci), dtype=torch.bfloat16, device=inp.device)
    rc = _L().vae_conv_umma_rw(_p(inp), _p(wpk), _f32(bias, "bias"), _p(out), _p(stats), B, Hi, Wi, Ci, Co,
                               stride, mode, _stream())
    _lib.check(rc, "vae_conv_umma_rw")
    return out


def conv_umma_act(kind: str, inp, wpk, bias, vec, beta, Ci: int, Co: int, stride: int, mode: int, slope: float = LRELU_SLOPE):
    """conv_umma_rw / conv_umma_tma with eval-mode BatchNorm + LeakyReLU in the epilogue (inference only): returns the
    block's ACTIVATION.  ``vec`` = the layer's (scale, shift, mean, rstd) rows from the running statistics."""
    if inp.dtype != torch.bfloat16:
        raise RuntimeError("conv_umma_act: bf16 activations only")
    B = inp.shape[0]
    if mode == 0:
        Hi, Wi = inp.shape[1], inp.shape[2]
        out = torch.empty((B, Hi // stride, Wi // stride, Co), dtype=torch.bfloat16, device=inp.device)
    else:
        Hi, Wi = inp.shape[1] * stride, inp.shape[2] * stride
        out = torch.empty((B, Hi, Wi, Ci), dtype=torch.bfloat16, device=inp.device)
    fn = _L().vae_conv_umma_rw_act if kind == "rw" else _L().vae_conv_umma_tma_act
    rc = fn(_p(inp), _p(wpk), _f32(bias, "bias"), _p(out), vec[2].data_ptr(), vec[0].data_ptr(), _f32(beta, "bn bias"), slope,
            B, Hi, Wi, Ci, Co, stride, mode, _stream())
    _lib.check(rc, f"vae_conv_umma_{kind}_act")
    return out


def conv_umma_bnbwd(kind: str, inp, wpk, y, vec, beta, Ci: int, Co: int, stride: int, mode: int, slope: float = LRELU_SLOPE):
    """conv_umma_rw / conv_umma_tma producing the gradient w.r.t. act = lrelu(bn(y)) with LeakyReLU' and the two
    BatchNorm-backward reductions fused into the epilogue -> (du, dstats[2C] fp64).  ``vec`` = the layer's
    (scale, shift, mean, rstd) rows, ``y`` its stored conv output (shape of the result)."""
    if inp.dtype != torch.bfloat16 or y.dtype != torch.bfloat16:
        raise RuntimeError("conv_umma_bnbwd: bf16 activations only")
    C = y.shape[-1]
    du = torch.empty_like(y)
    dstats = torch.empty(2 * C, dtype=torch.float64, device=y.device)
    B = inp.shape[0]
    if mode == 0:
        Hi, Wi = inp.shape[1], inp.shape[2]
    else:
        Hi, Wi = inp.shape[1] * stride, inp.shape[2] * stride
    fn = _L().vae_conv_umma_rw_bnbwd if kind == "rw" else _L().vae_conv_umma_tma_bnbwd
    rc = fn(_p(inp), _p(wpk), _p(du), _p(y), vec[2].data_ptr(), vec[0].data_ptr(), _f32(beta, "bn bias"), vec[3].data_ptr(),
            _p(dstats), slope, B, Hi, Wi, Ci, Co, stride, mode, _stream())
    _lib.check(rc, f"vae_conv_umma_{kind}_bnbwd")
    return du, dstats


class ConvWeights:
    """Per-step operand layouts of one 3x3 conv weight (fp32 OIHW master copy).

    bf16 layers whose shape suits the tensor cores get pre-swizzled tcgen05 weight
    images; everything else gets the CUDA-core K-major packs.  Layouts are built
    lazily, once per forward (the optimizer changes the master copy every step)."""

    __slots__ = ("w", "dtype", "stride", "Co", "Ci", "_simt", "_umma", "_rw", "_batch")

    def __init__(self, w: torch.Tensor, dtype: torch.dtype, stride: int):
        self.w, self.dtype, self.stride = w, dtype, stride
        self.Co, self.Ci = w.shape[0], w.shape[1]
        self._simt = None
        self._umma = {}
        self._rw = {}
        self._batch = None  # PackBatch whose side-stream launches fill the images held here

    def simt(self):
        if self._simt is None:
            self._simt = pack_conv_weight(self.w, self.dtype)
        return self._simt

    def rw(self, mode: int):
        if mode not in self._rw:
            self._rw[mode] = umma_rw_pack_weight(self.w, mode)
        elif self._batch is not None:
            self._batch.wait()
        return self._rw[mode]

    def umma(self, mode: int, tma: bool = False):
        if self.dtype != torch.bfloat16 or not umma_available():
            return None
        key = (mode, tma)
        if key not in self._umma:
            pack = umma_tma_pack_weight if tma else umma_pack_weight
            self._umma[key] = pack(self.w, self.stride, mode)
        elif self._batch is not None:
            self._batch.wait()
        return self._umma[key]

    def kind(self, B: int, Hi: int, Wi: int, mode: int) -> str:
        """Which tcgen05 kernel serves this layer at activation extent [B, Hi, Wi] (the high-resolution
        side): 'rw' (resident weights), 'tma' (streamed weights) or 'other'."""
        if self.dtype != torch.bfloat16 or not umma_available():
            return "other"
        if umma_rw_eligible(B, Hi, Wi, self.Ci, self.Co, self.stride, mode):
            return "rw"
        if umma_tma_eligible(B, Hi, Wi, self.Ci, self.Co, self.stride, mode):
            return "tma"
        return "other"


_AUX_STREAMS = {}


def aux_stream(device) -> "torch.cuda.Stream":
    idx = torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    if idx not in _AUX_STREAMS:
        _AUX_STREAMS[idx] = torch.cuda.Stream(device=idx)
    return _AUX_STREAMS[idx]


class ZeroArena:
    """Per-device scratch arena kept ZERO by its owner: the buffers kernels accumulate into (BatchNorm statistics, split-K
    weight-gradient workspaces, loss scratch) are carved from it and the library skips its own clear for pointers
    inside it (``vae_set_zero_arena``).  ``reset()`` -- called once at the start of every step by the step driver
    (trainer.TrainStep, bench.py) -- clears what the previous step handed out with ONE memset instead of ~25 memset
    nodes spread between the kernels of the step.  Everything past the cursor is zero at all times; a request that
    does not fit falls back to a fresh tensor (which the library clears as before)."""

    ALIGN = 256

    def __init__(self, device, nbytes: int):
        self.device = torch.device(device)
        self.buf = torch.zeros(int(nbytes), dtype=torch.uint8, device=self.device)
        self.off = 0
        with torch.cuda.device(self.device):
            _lib.check(_L().vae_set_zero_arena(self.buf.data_ptr(), self.buf.numel()), "vae_set_zero_arena")

    def take(self, shape, dtype):
        n = int(math.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        if self.off + n > self.buf.numel():
            return None
        t = self.buf[self.off : self.off + n].view(dtype).view(shape)
        self.off += (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        return t

    def reset(self) -> None:
        if self.off:
            self.buf[: min(self.off, self.buf.numel())].zero_()
            self.off = 0

    def close(self) -> None:
        with torch.cuda.device(self.device):
            _L().vae_set_zero_arena(None, 0)


_ZERO_ARENAS = {}


def _dev_index(device) -> int:
    idx = torch.device(device).index if device is not None else None
    return torch.cuda.current_device() if idx is None else idx


def enable_zero_arena(device=None, nbytes: int = 64 << 20) -> ZeroArena:
    """Turn the scratch arena on for ``device``.  Only for callers that run whole steps and call
    ``zero_arena_reset`` at the start of each one (scratch handed out in a step must be dead by the next reset)."""
    idx = _dev_index(device)
    if idx not in _ZERO_ARENAS:
        _ZERO_ARENAS[idx] = ZeroArena(torch.device("cuda", idx), nbytes)
    return _ZERO_ARENAS[idx]


def disable_zero_arena(device=None) -> None:
    a = _ZERO_ARENAS.pop(_dev_index(device), None)
    if a is not None:
        a.close()


def zero_arena_reset(device=None) -> None:
    a = _ZERO_ARENAS.get(_dev_index(device))
    if a is not None:
        a.reset()


def scratch(shape, dtype, device):
    """Buffer a kernel accumulates into (the library clears it first -- or the arena already did)."""
    shape = (shape,) if isinstance(shape, int) else tuple(shape)
    a = _ZERO_ARENAS.get(_dev_index(device)) if _ZERO_ARENAS else None
    if a is not None:
        t = a.take(shape, dtype)
        if t is not None:
            return t
    return torch.empty(shape, dtype=dtype, device=device)


_COPY_STREAMS = {}


def copy_stream(device) -> "torch.cuda.Stream":
    """Side stream for device->host result copies that overlap the next chunk's kernels (VanillaVAE.sample_to_host)."""
    idx = torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    if idx not in _COPY_STREAMS:
        _COPY_STREAMS[idx] = torch.cuda.Stream(device=idx)
    return _COPY_STREAMS[idx]


class PackBatch:
    """All weight images of one step, packed on a side stream while the first (CUDA-core) layer runs.

    The images depend only on the fp32 master weights, so their ~15 small launches need not sit on the
    critical path between the layers.  Buffers are allocated on the caller's stream (they are freed
    there), the side stream is forked from it (the optimizer step of the previous iteration precedes the
    fork) and joined by ``wait()`` at the first use -- a fork/join pattern CUDA-graph capture accepts."""

    def __init__(self):
        self._jobs = []
        self._event = None

    def add(self, cw: ConvWeights, B: int, Hi: int, Wi: int, mode: int) -> None:
        kind = cw.kind(B, Hi, Wi, mode)
        if kind == "rw" and mode not in cw._rw:
            out = torch.empty(9 * cw.Ci * cw.Co, dtype=torch.bfloat16, device=cw.w.device)
            cw._rw[mode] = out
            self._jobs.append(lambda: umma_rw_pack_weight(cw.w, mode, out=out))
        elif kind == "tma" and (mode, True) not in cw._umma:
            n = int(_L().vae_umma_tma_packed_elems(cw.Ci, cw.Co, cw.stride, mode))
            if n == 0:
                return
            out = torch.empty(n, dtype=torch.bfloat16, device=cw.w.device)
            cw._umma[(mode, True)] = out
            self._jobs.append(lambda: umma_tma_pack_weight(cw.w, cw.stride, mode, out=out))
        else:
            return
        cw._batch = self

    def launch(self) -> None:
        if not self._jobs:
            return
        dev = torch.cuda.current_device()
        main = torch.cuda.current_stream(dev)
        aux = aux_stream(dev)
        aux.wait_stream(main)
        with torch.cuda.stream(aux):
            for job in self._jobs:
                job()
            self._event = torch.cuda.Event()
            self._event.record(aux)
        self._jobs = []

    def wait(self) -> None:
        if self._event is not None:
            torch.cuda.current_stream().wait_event(self._event)
            self._event = None


def conv_forward(x, cw: ConvWeights, bias, stats=None, Ci: Optional[int] = None, pro=None, bnbwd=None, act=None):
    """Conv2d(k3, stride, p1) forward: tcgen05 kernel when the layer qualifies.

    ``bnbwd`` = (y, vec, beta) of the BatchNorm block whose activation gradient this call produces (the conv runs
    as the backward of a ConvTranspose2d): when a tcgen05 kernel serves the layer the result is (du, dstats) with
    LeakyReLU' and the BatchNorm-backward reductions fused into the epilogue, else (d_act, None)."""
    if act is not None:   # inference: (vec, beta) of the block's eval-mode BatchNorm -> the result is the ACTIVATION
        vec_, beta_ = act
        if x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[-1] == cw.Ci and pro is None and umma_available():
            if umma_rw_eligible(x.shape[0], x.shape[1], x.shape[2], cw.Ci, cw.Co, cw.stride, 0):
                return conv_umma_act("rw", x, cw.rw(0), bias, vec_, beta_, cw.Ci, cw.Co, cw.stride, 0)
            if umma_tma_eligible(x.shape[0], x.shape[1], x.shape[2], cw.Ci, cw.Co, cw.stride, 0):
                return conv_umma_act("tma", x, cw.umma(0, tma=True), bias, vec_, beta_, cw.Ci, cw.Co, cw.stride, 0)
        return bn_apply_lrelu(conv_forward(x, cw, bias, None, Ci, pro), vec_, beta_)
    if bnbwd is not None:
        y_, vec_, beta_ = bnbwd
        if x.dtype == torch.bfloat16 and pro is None and bias is None and umma_available() and y_.dtype == torch.bfloat16:
            if umma_rw_eligible(x.shape[0], x.shape[1], x.shape[2], cw.Ci, cw.Co, cw.stride, 0):
                return conv_umma_bnbwd("rw", x, cw.rw(0), y_, vec_, beta_, cw.Ci, cw.Co, cw.stride, 0)
            if umma_tma_eligible(x.shape[0], x.shape[1], x.shape[2], cw.Ci, cw.Co, cw.stride, 0):
                return conv_umma_bnbwd("tma", x, cw.umma(0, tma=True), y_, vec_, beta_, cw.Ci, cw.Co, cw.stride, 0)
        return conv_forward(x, cw, bias, stats, Ci, pro), None
    if x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[-1] == cw.Ci:
        if pro is None and umma_available() and umma_rw_eligible(x.shape[0], x.shape[1], x.shape[2], cw.Ci, cw.Co, cw.stride, 0):
            return conv_umma_rw(x, cw.rw(0), bias, stats, cw.Ci, cw.Co, cw.stride, 0)
        if pro is None and umma_tma_eligible(x.shape[0], x.shape[1], x.shape[2], cw.Ci, cw.Co, cw.stride, 0):
            return conv_umma_tma(x, cw.umma(0, tma=True), bias, stats, cw.Ci, cw.Co, cw.stride, 0)
        wpk = cw.umma(0)
        if wpk is not None:
            return conv_umma(x, wpk, bias, stats, cw.Ci, cw.Co, cw.stride, 0, pro=pro)
    if pro is not None:
        raise RuntimeError("fused BatchNorm prologue needs the tcgen05 path")
    return conv_fprop(x, cw.simt()[0], bias, cw.stride, cw.dtype, stats, Ci=Ci if Ci is not None else cw.Ci)


def conv_data_grad(dy, cw: ConvWeights, bias=None, stats=None, pro=None, bnbwd=None, act=None):
    """Data gradient of that conv == ConvTranspose2d(k3, stride, p1, op=stride-1) forward.
    ``bnbwd``: see conv_forward."""
    if act is not None:   # inference: see conv_forward
        vec_, beta_ = act
        s_ = cw.stride
        if dy.dtype == torch.bfloat16 and pro is None and umma_available():
            if umma_rw_eligible(dy.shape[0], dy.shape[1] * s_, dy.shape[2] * s_, cw.Ci, cw.Co, s_, 1):
                return conv_umma_act("rw", dy, cw.rw(1), bias, vec_, beta_, cw.Ci, cw.Co, s_, 1)
            if umma_tma_eligible(dy.shape[0], dy.shape[1] * s_, dy.shape[2] * s_, cw.Ci, cw.Co, s_, 1):
                return conv_umma_act("tma", dy, cw.umma(1, tma=True), bias, vec_, beta_, cw.Ci, cw.Co, s_, 1)
        return bn_apply_lrelu(conv_data_grad(dy, cw, bias, None, pro), vec_, beta_)
    if bnbwd is not None:
        y_, vec_, beta_ = bnbwd
        s_ = cw.stride
        if dy.dtype == torch.bfloat16 and pro is None and bias is None and umma_available() and y_.dtype == torch.bfloat16:
            if umma_rw_eligible(dy.shape[0], dy.shape[1] * s_, dy.shape[2] * s_, cw.Ci, cw.Co, s_, 1):
                return conv_umma_bnbwd("rw", dy, cw.rw(1), y_, vec_, beta_, cw.Ci, cw.Co, s_, 1)
            if umma_tma_eligible(dy.shape[0], dy.shape[1] * s_, dy.shape[2] * s_, cw.Ci, cw.Co, s_, 1):
                return conv_umma_bnbwd("tma", dy, cw.umma(1, tma=True), y_, vec_, beta_, cw.Ci, cw.Co, s_, 1)
        return conv_data_grad(dy, cw, bias, stats, pro), None
    if dy.dtype == torch.bfloat16:
        s_ = cw.stride
        if pro is None and umma_available() and umma_rw_eligible(dy.shape[0], dy.shape[1] * s_, dy.shape[2] * s_, cw.Ci, cw.Co, s_, 1):
            return conv_umma_rw(dy, cw.rw(1), bias, stats, cw.Ci, cw.Co, s_, 1)
        if pro is None and umma_tma_eligible(dy.shape[0], dy.shape[1] * s_, dy.shape[2] * s_, cw.Ci, cw.Co, s_, 1):
            return conv_umma_tma(dy, cw.umma(1, tma=True), bias, stats, cw.Ci, cw.Co, s_, 1)
        wpk = cw.umma(1)
        if wpk is not None:
            return conv_umma(dy, wpk, bias, stats, cw.Ci, cw.Co, cw.stride, 1, pro=pro)
    if pro is not None:
        raise RuntimeError("fused BatchNorm prologue needs the tcgen05 path")
    return conv_dgrad(dy, cw.simt()[1], bias, cw.stride, stats)


def conv_weight_grad(x, dy, stride: int, Ci: int, Co: int, Hi: int, Wi: int, out=None):
    """dW [Co,Ci,3,3] fp32: tcgen05 kernel for bf16 tensor-core shaped layers, the
    CUDA-core kernels otherwise (fp32 path, single-channel edge layers).  ``out``: optional
    destination in the reference's OIHW layout (a gradient-bucket view)."""
    if out is not None and (out.dtype != torch.float32 or out.numel() != Co * Ci * 9):
        raise RuntimeError("conv_weight_grad: destination must be float32 [Co, Ci, 3, 3]")
    if x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16 and umma_available() and \
            _L().vae_wgrad_tma_eligible(dy.shape[0], Hi, Wi, Ci, Co, stride):
        B = dy.shape[0]
        dw = out if out is not None else torch.empty((Co, Ci, 3, 3), dtype=torch.float32, device=dy.device)
        ws = scratch(9 * Co * Ci, torch.float32, dy.device)
        rc = _L().vae_conv_wgrad_tma(_p(x), _p(dy), _p(dw), _p(ws), B, Hi, Wi, Ci, Co, stride, _stream())
        _lib.check(rc, "vae_conv_wgrad_tma")
        return dw
    if x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16 and _L().vae_umma_wgrad_supported(Ci, Co):
        B = dy.shape[0]
        dw = out if out is not None else torch.empty((Co, Ci, 3, 3), dtype=torch.float32, device=dy.device)
        rc = _L().vae_conv_wgrad_umma(_p(x), _p(dy), _p(dw), B, Hi, Wi, Ci, Co, stride, _stream())
        _lib.check(rc, "vae_conv_wgrad_umma")
        return dw
    return conv_wgrad(x, dy, stride, Ci, Co, Hi, Wi, out=out)

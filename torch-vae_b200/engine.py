"""Forward / backward schedules of the VAE hot path as autograd Functions.

The reference runs ``models.py:185-188`` (forward), ``:190-225`` (loss) and
autograd's backward (``train.py:650``) as ~120 ATen ops.  Here each of the four
pieces -- encode, reparameterize, decode, ELBO -- is ONE autograd node whose
forward and backward are explicit schedules of the sm_100a kernels behind
``include/vae_b200.h``:

  activations NHWC in the compute dtype (fp32 or bf16), accumulators fp32,
  BatchNorm statistics fp64, parameters and their gradients fp32 in the
  reference layouts (OIHW / [out,in]).

ConvTranspose2d(k3,s2,p1,op1) forward IS the data-gradient of the equivalent
Conv2d(k3,s2,p1), so three kernel families (fprop, dgrad,
wgrad) cover all nine conv layers in both directions.
"""

from __future__ import annotations

from typing import List, Optional

import torch
from torch.autograd.function import once_differentiable

from . import ops


class ConvBlockSpec:
    """Static description of one conv(+BN+LeakyReLU) block."""

    __slots__ = ("transposed", "ci", "co", "bn")

    def __init__(self, transposed: bool, ci: int, co: int, bn):
        self.transposed = transposed  # True: ConvTranspose2d (runs the dgrad kernel forward)
        self.ci = ci  # channels on the HIGH-resolution side of the equivalent conv
        self.co = co  # channels on the LOW-resolution side
        self.bn = bn  # nn.BatchNorm2d holding running stats (buffers are not autograd inputs)


def _bn_vectors(y, stats, gamma, beta, bn, training: bool):
    """BatchNorm statistics -> (scale, shift, mean, rstd) + running-stat update; no apply pass."""
    count = y.numel() // y.shape[-1]
    use_batch = training or bn.running_mean is None
    vec = ops.bn_finalize(
        stats if use_batch else None, gamma, beta, bn.running_mean, bn.running_var, count, use_batch,
        momentum=bn.momentum if bn.momentum is not None else 0.1, eps=bn.eps,
    )
    if training and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    return vec


def _bn_forward(y, stats, gamma, beta, bn, training: bool):
    count = y.numel() // y.shape[-1]
    C = y.shape[-1]
    if C % 8 == 0 and C // 8 <= 256:
        use_batch = training or bn.running_mean is None
        nbt = bn.num_batches_tracked if (training and bn.num_batches_tracked is not None) else None
        return ops.bn_finalize_apply_lrelu(
            y, stats if use_batch else None, gamma, beta, bn.running_mean, bn.running_var, nbt, count, use_batch,
            momentum=bn.momentum if bn.momentum is not None else 0.1, eps=bn.eps)
    use_batch = training or bn.running_mean is None
    vec = ops.bn_finalize(
        stats if use_batch else None, gamma, beta, bn.running_mean, bn.running_var, count, use_batch,
        momentum=bn.momentum if bn.momentum is not None else 0.1, eps=bn.eps,
    )
    if training and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    return vec, ops.bn_apply_lrelu(y, vec, beta)


# LeakyReLU' + the two BatchNorm-backward reductions fused into the epilogue of the conv that produces the
# activation gradient (vae_conv_umma_{rw,tma}_bnbwd).  Correct (tests/test_umma_gpu.py, teacher-forced hops) but
# measured SLOWER than the separate reduction pass on the B200 (the 8 epilogue warps of the
# persistent conv kernels are issue / latency bound), so the step does not use it.
FUSE_BN_BACKWARD = False

# Inference (model.eval() and nothing requires a gradient: sample(), evaluate()): eval-mode BatchNorm + LeakyReLU are
# applied in the epilogue of the conv that produces the block (vae_conv_umma_{rw,tma}_act), so ConvTranspose2d + BN +
# LeakyReLU is ONE pass and the pre-normalisation tensor is never written.
INFERENCE_EPILOGUE = True


def inference_mode(module_training: bool, *tensors) -> bool:
    """True when the call is a pure inference: eval mode and autograd will record nothing (no_grad, or no input /
    parameter requires a gradient).  Evaluated by the caller -- inside Function.forward grad mode is always off."""
    if module_training:
        return False
    return not (torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors))


def _bnbwd(saved_prev, beta_prev):
    return (saved_prev[1], saved_prev[2], beta_prev) if FUSE_BN_BACKWARD else None


def _conv_bnbwd(fn, dy, cw, bnbwd):
    """fn(dy, cw[, bnbwd]) -> (d_act or du, dstats or None)."""
    if bnbwd is None:
        return (fn(dy, cw) if fn is ops.conv_data_grad else fn(dy, cw, None)), None
    return fn(dy, cw, bnbwd=bnbwd) if fn is ops.conv_data_grad else fn(dy, cw, None, bnbwd=bnbwd)


# Teacher-forcing hook for the parity tests: when a list is installed here, every backward hop appends
# (name, {tensor name: tensor}) with exactly the tensors its kernels consumed and produced, so that a test can
# redo that hop in exact arithmetic on IDENTICAL inputs (tests/test_bf16_rigor_gpu.py).  None in production.
TRACE = None


def _trace(name, **tensors):
    if TRACE is not None:
        TRACE.append((name, tensors))


def _dst(sink, p):
    """Where a parameter gradient is written: straight into its all-reduce bucket slice when a
    gradient sink (parallel.GradientBuckets) is attached, else a fresh tensor (None)."""
    return sink.view(p) if sink is not None else None


def _dst3(sink, params):
    return tuple(sink.view(p).view(-1) for p in params) if sink is not None else None


class EncodeFn(torch.autograd.Function):
    """encoder convs -> flatten -> fc_mu / fc_var  (models.py:107-145)."""

    @staticmethod
    def forward(ctx, specs: List[ConvBlockSpec], dtype, training, infer, sink, cws, x, w_mu, b_mu, w_var, b_var, *conv_params):
        ctx.set_materialize_grads(False)
        ctx.sink = sink
        infer = bool(infer) and INFERENCE_EPILOGUE and not training   # eval mode and no backward will follow (see inference_mode)
        B, Cin = x.shape[0], x.shape[1]
        h = x.contiguous()
        if h.dtype != torch.float32:
            h = h.float()
        Hi, Wi = h.shape[2], h.shape[3]
        if Cin > 1:  # NCHW image -> NHWC; with one channel the two layouts coincide
            h = ops.nchw_to_nhwc(h.view(B, -1), Cin, Hi, Wi)
        saved = []
        for i, sp in enumerate(specs):
            w, b, gamma, beta = conv_params[4 * i : 4 * i + 4]
            cw = cws[i] if cws is not None else ops.ConvWeights(w, dtype, 2)
            if infer and sp.bn.running_mean is not None and h.dtype == torch.bfloat16:
                # eval-mode BatchNorm + LeakyReLU ride in the conv's epilogue: the block writes its activation
                vec = _bn_vectors(h, None, gamma, beta, sp.bn, False)
                h = ops.conv_forward(h, cw, b, None, Ci=sp.ci, act=(vec, beta))
                Hi, Wi = Hi // 2, Wi // 2
                continue
            stats = ops.scratch(2 * sp.co, torch.float64, h.device) if training else None
            y = ops.conv_forward(h, cw, b, stats, Ci=sp.ci)
            vec, act = _bn_forward(y, stats, gamma, beta, sp.bn, training)
            saved.append((h, y, vec, cw, Hi, Wi))
            h = act
            Hi, Wi = Hi // 2, Wi // 2
        pre = ops.nhwc_to_nchw(h)  # [B, C*h*w], C-major like x.flatten(1) (models.py:133)
        F_, L_ = pre.shape[1], w_mu.shape[0]
        if ops.heads_supported(F_, L_):
            mu, log_var = ops.skinny_reduce_pair(pre, w_mu, b_mu, w_var, b_var, F_, 1, L_)   # `pre` read once
        else:
            mu = ops.linear_fwd(pre, w_mu, b_mu, torch.float32)
            log_var = ops.linear_fwd(pre, w_var, b_var, torch.float32)
        ctx.specs, ctx.training, ctx.saved, ctx.feat = specs, training, saved, (specs[-1].co, Hi, Wi)
        ctx.save_for_backward(pre, w_mu, w_var, b_mu, b_var, *conv_params)
        ctx.x_needs_grad = x.requires_grad
        return mu, log_var, pre

    @staticmethod
    @once_differentiable
    def backward(ctx, dmu, dlv, dpre):
        pre, w_mu, w_var, b_mu, b_var, *conv_params = ctx.saved_tensors
        specs, saved, sink = ctx.specs, ctx.saved, ctx.sink
        dev = pre.device
        B, L = pre.shape[0], w_mu.shape[0]
        if dmu is None:
            dmu = torch.zeros((B, L), dtype=torch.float32, device=dev)
        if dlv is None:
            dlv = torch.zeros((B, L), dtype=torch.float32, device=dev)
        dmu, dlv = dmu.contiguous(), dlv.contiguous()
        F_ = pre.shape[1]
        if ops.heads_supported(F_, L):
            d_pre = ops.skinny_expand(dmu, w_mu, dlv, w_var, None, F_, 1, F_, pre.dtype)
            g_w_mu, _ = ops.skinny_wgrad(dmu, pre, (L, F_), F_, 1, False, out=_dst(sink, w_mu))
            g_w_var, _ = ops.skinny_wgrad(dlv, pre, (L, F_), F_, 1, False, out=_dst(sink, w_var))
        else:
            d_pre = ops.linear_bwd_input_pair(dmu, w_mu, dlv, w_var, pre.dtype)
            g_w_mu = ops.linear_bwd_weight(dmu, pre, out=_dst(sink, w_mu))
            g_w_var = ops.linear_bwd_weight(dlv, pre, out=_dst(sink, w_var))
        if dpre is not None:
            d_pre = d_pre + dpre.to(d_pre.dtype)
        g_b_mu, g_b_var = ops.colsum(dmu, out=_dst(sink, b_mu)), ops.colsum(dlv, out=_dst(sink, b_var))
        _trace("heads", dmu=dmu, dlv=dlv, pre=pre, w_mu=w_mu, w_var=w_var, d_pre=d_pre, g_w_mu=g_w_mu, g_w_var=g_w_var,
               g_b_mu=g_b_mu, g_b_var=g_b_var)
        if sink is not None:
            for p_, g_ in ((w_mu, g_w_mu), (b_mu, g_b_mu), (w_var, g_w_var), (b_var, g_b_var)):
                sink.push(p_, g_)
        C, h, w = ctx.feat
        d_act = ops.nchw_to_nhwc(d_pre, C, h, w)
        conv_grads: List[Optional[torch.Tensor]] = [None] * len(conv_params)
        dx = None
        dstats = None   # set when d_act already holds du: the producing conv fused LeakyReLU' + the BN reductions
        for i in range(len(specs) - 1, -1, -1):
            sp = specs[i]
            h_in, y, vec, cw, Hi, Wi = saved[i]
            gamma, beta = conv_params[4 * i + 2], conv_params[4 * i + 3]
            w_i, b_i = conv_params[4 * i], conv_params[4 * i + 1]
            dy, dgamma, dbeta, dbias = ops.bn_lrelu_backward(d_act, y, vec, gamma, beta, ctx.training,
                                                             outs=_dst3(sink, (gamma, beta, b_i)), dstats=dstats)
            conv_grads[4 * i] = ops.conv_weight_grad(h_in, dy, 2, sp.ci, sp.co, Hi, Wi, out=_dst(sink, w_i))
            conv_grads[4 * i + 1] = dbias
            conv_grads[4 * i + 2] = dgamma
            conv_grads[4 * i + 3] = dbeta
            if sink is not None:
                for k in range(4):
                    sink.push(conv_params[4 * i + k], conv_grads[4 * i + k])
            d_in, d_in_is_du = d_act, dstats is not None
            if i > 0:
                # gradient w.r.t. the activation of block i-1, with that block's LeakyReLU' and BN reductions fused
                d_act, dstats = _conv_bnbwd(ops.conv_data_grad, dy, cw, _bnbwd(saved[i - 1], conv_params[4 * i - 1]))
            elif ctx.x_needs_grad:
                dx = ops.nhwc_to_nchw(ops.conv_data_grad(dy, cw), torch.float32).view(B, sp.ci, Hi, Wi)
            _trace(f"encoder.{i}", transposed=False, d_act=d_in, d_act_is_du=d_in_is_du, y=y, vec=vec, gamma=gamma, h_in=h_in,
                   w=conv_params[4 * i], dy=dy, dgamma=dgamma, dbeta=dbeta, dw=conv_grads[4 * i], d_out=d_act if i > 0 else None,
                   d_out_bn=(saved[i - 1][1], saved[i - 1][2]) if i > 0 and dstats is not None else None,
                   training=ctx.training)
        ctx.saved = None
        if sink is not None:
            # the sink owns the parameter gradients (it exposes them as p.grad after the all-reduce); handing
            # them to autograd as well would let AccumulateGrad add into a bucket that is being reduced
            return (None, None, None, None, None, None, dx, None, None, None, None, *([None] * len(conv_grads)))
        return (None, None, None, None, None, None, dx, g_w_mu, g_b_mu, g_w_var, g_b_var, *conv_grads)


class ReparamFn(torch.autograd.Function):
    """z = mu + eps * exp(0.5 * log_var)  (models.py:177-183)."""

    @staticmethod
    def forward(ctx, mu, log_var, eps, seed, dev_offset):
        mu, log_var = mu.contiguous(), log_var.contiguous()
        if eps is not None:
            eps = eps.to(device=mu.device, dtype=torch.float32).contiguous()
        z, eps_used = ops.reparam_fwd(mu, log_var, eps, seed=seed, offset=0, dev_offset=dev_offset)
        if dev_offset is not None and eps is None:
            dev_offset.add_((mu.numel() + 3) // 4)  # advance the Philox counter (graph-capturable)
        ctx.save_for_backward(eps_used, log_var)
        return z

    @staticmethod
    @once_differentiable
    def backward(ctx, dz):
        eps, log_var = ctx.saved_tensors
        dmu, dlv = ops.reparam_bwd(dz.contiguous(), eps, log_var)
        _trace("reparam", dz=dz, eps=eps, log_var=log_var, dmu=dmu, dlv=dlv)
        return dmu, dlv, None, None, None


class DecodeFn(torch.autograd.Function):
    """decoder_input -> view -> ConvT blocks -> final ConvT/BN/LReLU/Conv -> sigmoid
    (models.py:147-175).  Returns (probabilities fp32, logits in compute dtype)."""

    @staticmethod
    def forward(ctx, specs: List[ConvBlockSpec], feat, dtype, training, infer, sink, cws, z, w_in, b_in, w_last, b_last, *conv_params):
        ctx.set_materialize_grads(False)
        ctx.sink = sink
        infer = bool(infer) and INFERENCE_EPILOGUE and not training   # eval mode and no backward will follow (see inference_mode)
        z = z.contiguous()
        if z.dtype != torch.float32:
            z = z.float()
        B = z.shape[0]
        C, h0, w0 = feat
        L_, F_ = z.shape[1], w_in.shape[0]
        if ops.heads_supported(F_, L_):
            d = ops.skinny_expand(z, w_in, None, None, b_in, 1, L_, F_, dtype)  # [B, C*h*w], C-major
        else:
            d = ops.linear_fwd(z, w_in, b_in, dtype)
        h = ops.nchw_to_nhwc(d, C, h0, w0)
        saved = []
        fused_last = False
        for i, sp in enumerate(specs):
            w, b, gamma, beta = conv_params[4 * i : 4 * i + 4]
            cw = cws[i] if cws is not None else ops.ConvWeights(w, dtype, 2)
            last_fused = i == len(specs) - 1 and ops.final_block_supported(B, h.shape[1] * 2, h.shape[2] * 2, sp.ci)
            if infer and not last_fused and sp.bn.running_mean is not None and h.dtype == torch.bfloat16:
                vec = _bn_vectors(h, None, gamma, beta, sp.bn, False)
                h = ops.conv_data_grad(h, cw, b, None, act=(vec, beta))  # ConvTranspose2d + eval BatchNorm + LeakyReLU, one pass
                continue
            stats = ops.scratch(2 * sp.ci, torch.float64, z.device) if training else None
            y = ops.conv_data_grad(h, cw, b, stats)  # ConvTranspose2d forward
            if i == len(specs) - 1 and ops.final_block_supported(B, y.shape[1], y.shape[2], sp.ci):
                # output block: BatchNorm apply + LeakyReLU ride in the final conv's operand load
                vec = _bn_vectors(y, stats, gamma, beta, sp.bn, training)
                saved.append((h, y, vec, cw))
                fused_last = True
                break
            vec, act = _bn_forward(y, stats, gamma, beta, sp.bn, training)
            saved.append((h, y, vec, cw))
            h = act
        if fused_last:
            logits, probs = ops.final_conv_fwd(y, vec, beta, w_last, b_last)
            cw3 = None
        else:
            cw3 = ops.ConvWeights(w_last, dtype, 1)
            logits = ops.conv_forward(h, cw3, b_last)  # [B,H,W,1]
            H, W = logits.shape[1], logits.shape[2]
            probs = ops.sigmoid(logits).view(B, 1, H, W)
            logits = logits.view(B, 1, H, W)
        ctx.specs, ctx.training, ctx.saved, ctx.last = specs, training, saved, (None if fused_last else h, cw3)
        ctx.save_for_backward(z, w_in, probs, b_in, w_last, b_last, *conv_params)
        ctx.z_needs_grad = z.requires_grad
        return probs, logits

    @staticmethod
    @once_differentiable
    def backward(ctx, dprobs, dlogits):
        z, w_in, probs, b_in, w_last, b_last, *conv_params = ctx.saved_tensors
        specs, saved, sink = ctx.specs, ctx.saved, ctx.sink
        act_last, cw3 = ctx.last
        B, _, H, W = probs.shape
        fused_last = cw3 is None
        dtype = saved[-1][1].dtype
        if dlogits is not None:
            dlog = dlogits.contiguous()
            if dprobs is not None:
                dlog = ops.sigmoid_bwd(dprobs.contiguous(), probs, dlogits=dlog.clone())
        elif dprobs is not None:
            dlog = ops.sigmoid_bwd(dprobs.contiguous(), probs, dtype=dtype)
        else:
            dlog = torch.zeros((B, 1, H, W), dtype=dtype, device=probs.device)
        dlog = dlog.view(B, H, W, 1)
        conv_grads: List[Optional[torch.Tensor]] = [None] * len(conv_params)
        n_blk = len(specs)
        dstats = None   # see EncodeFn.backward
        if fused_last:
            # Conv2d(C->1) backward + LeakyReLU + BatchNorm backward in two passes over y
            _, y, vec, _ = saved[-1]
            gamma, beta = conv_params[4 * (n_blk - 1) + 2], conv_params[4 * (n_blk - 1) + 3]
            b_conv = conv_params[4 * (n_blk - 1) + 1]
            fused = ops.final_block_backward(dlog, y, vec, gamma, beta, w_last, ctx.training,
                                             outs=_dst3(sink, (gamma, beta, b_conv)),
                                             last_outs=(sink.view(w_last), sink.view(b_last)) if sink is not None else None)
            g_w_last, g_b_last = fused[4], fused[5]
            d_act = None
            _trace("final", dlog=dlog, y=y, vec=vec, gamma=gamma, w_last=w_last, dy=fused[0], dgamma=fused[1], dbeta=fused[2],
                   g_w_last=g_w_last, g_b_last=g_b_last, training=ctx.training)
        else:
            c_last = act_last.shape[-1]
            # final Conv2d(C -> 1): bias grad is a plain sum, done as two column sums
            n = dlog.numel()
            cols = 256 if n % 256 == 0 else 1
            g_b_last = ops.colsum(ops.colsum(dlog.view(n // cols, cols)).view(cols, 1), out=_dst(sink, b_last))
            g_w_last = ops.conv_weight_grad(act_last, dlog, 1, c_last, 1, H, W, out=_dst(sink, w_last))
            d_act = ops.conv_data_grad(dlog, cw3)
        if sink is not None:
            sink.push(w_last, g_w_last)
            sink.push(b_last, g_b_last)
        for i in range(n_blk - 1, -1, -1):
            sp = specs[i]
            h_in, y, vec, cw = saved[i]
            gamma, beta = conv_params[4 * i + 2], conv_params[4 * i + 3]
            if fused_last and i == n_blk - 1:
                dy, dgamma, dbeta, dbias = fused[:4]
            else:
                dy, dgamma, dbeta, dbias = ops.bn_lrelu_backward(d_act, y, vec, gamma, beta, ctx.training,
                                                                 outs=_dst3(sink, (gamma, beta, conv_params[4 * i + 1])),
                                                                 dstats=dstats)
            Hi, Wi = y.shape[1], y.shape[2]
            # equivalent conv: X = convT output side (dy), Y = convT input (h_in)
            conv_grads[4 * i] = ops.conv_weight_grad(dy, h_in, 2, sp.ci, sp.co, Hi, Wi, out=_dst(sink, conv_params[4 * i]))
            conv_grads[4 * i + 1] = dbias
            conv_grads[4 * i + 2] = dgamma
            conv_grads[4 * i + 3] = dbeta
            if sink is not None:
                for k in range(4):
                    sink.push(conv_params[4 * i + k], conv_grads[4 * i + k])
            d_in, d_in_is_du = d_act, dstats is not None
            if i > 0:   # gradient w.r.t. the ConvT input = activation of block i-1: its LeakyReLU' + BN reductions fused
                d_act, dstats = _conv_bnbwd(ops.conv_forward, dy, cw, _bnbwd(saved[i - 1], conv_params[4 * i - 1]))
            else:       # decoder_input's output has no BatchNorm
                d_act, dstats = ops.conv_forward(dy, cw, None), None
            _trace(f"decoder.{i}", transposed=True, d_act=d_in, d_act_is_du=d_in_is_du, y=y, vec=vec, gamma=gamma, h_in=h_in,
                   w=conv_params[4 * i], dy=dy, dgamma=dgamma, dbeta=dbeta, dw=conv_grads[4 * i], d_out=d_act,
                   d_out_bn=(saved[i - 1][1], saved[i - 1][2]) if i > 0 and dstats is not None else None,
                   training=ctx.training, fused_bn=bool(fused_last and i == n_blk - 1))
        d_d = ops.nhwc_to_nchw(d_act)  # [B, F] C-major, matching decoder_input's output
        L_, F_ = z.shape[1], w_in.shape[0]
        skinny = ops.heads_supported(F_, L_)
        if skinny:
            g_w_in, g_b_in = ops.skinny_wgrad(z, d_d, (F_, L_), 1, L_, True, out=_dst(sink, w_in), colsum_out=_dst(sink, b_in))
        else:
            g_w_in = ops.linear_bwd_weight(d_d, z, out=_dst(sink, w_in))
            g_b_in = ops.colsum(d_d, out=_dst(sink, b_in))
        if sink is not None:
            sink.push(w_in, g_w_in)
            sink.push(b_in, g_b_in)
        if not ctx.z_needs_grad:
            dz = None
        elif skinny:
            dz = ops.skinny_reduce(d_d, w_in, None, 1, L_, L_)
        else:
            dz = ops.linear_bwd_input(d_d, w_in, out_dtype=torch.float32)
        _trace("decoder_input", d_d=d_d, z=z, w_in=w_in, g_w_in=g_w_in, g_b_in=g_b_in, dz=dz)
        ctx.saved = None
        ctx.last = None
        if sink is not None:  # see EncodeFn.backward: the sink owns the parameter gradients
            return (None, None, None, None, None, None, None, dz, None, None, None, None, *([None] * len(conv_grads)))
        return (None, None, None, None, None, None, None, dz, g_w_in, g_b_in, g_w_last, g_b_last, *conv_grads)


class ElboFn(torch.autograd.Function):
    """loss = mean BCE(sigmoid(logits), x) + kld_weight * KL  (models.py:190-225).

    When a gradient will be asked for, forward runs the one-pass kernel that also writes
    dlogits / dmu / dlogvar (the loss is the root of the graph); the first backward only
    rescales them by the upstream gradient (a no-op launch for ``loss.backward()``)."""

    @staticmethod
    def forward(ctx, pred, target, mu, log_var, kld_weight: float, input_is_prob: bool):
        pred = pred.contiguous()
        target = target.contiguous()
        if target.dtype != torch.float32:
            target = target.float()
        mu, log_var = mu.contiguous(), log_var.contiguous()
        fp32_latents = mu.dtype == torch.float32 and log_var.dtype == torch.float32
        if fp32_latents and (ctx.needs_input_grad[0] or ctx.needs_input_grad[2] or ctx.needs_input_grad[3]):
            losses, dpred, dmu, dlv = ops.elbo_fused(pred, target, mu, log_var, kld_weight, input_is_prob)
            ctx.cached = (dpred, dmu, dlv)
        else:
            losses = ops.elbo_fwd(pred, target, mu, log_var, kld_weight, input_is_prob)
            ctx.cached = None
        ctx.save_for_backward(pred, target, mu, log_var)
        ctx.kld_weight, ctx.input_is_prob = kld_weight, input_is_prob
        ctx.mark_non_differentiable(losses)
        return losses[0].clone(), losses

    @staticmethod
    @once_differentiable
    def backward(ctx, g, _unused):
        g = g.to(torch.float32).contiguous()
        cached, ctx.cached = ctx.cached, None
        if cached is not None:
            dpred, dmu, dlv = cached
            ops.elbo_scale_grads(g, dpred, dmu, dlv)
        else:  # a second backward through a retained graph, or latents that were not fp32
            pred, target, mu, log_var = ctx.saved_tensors
            dpred, dmu, dlv = ops.elbo_bwd(pred, target, mu, log_var, g, ctx.kld_weight, ctx.input_is_prob)
        if TRACE is not None:
            pred, target, mu, log_var = ctx.saved_tensors
            _trace("elbo", pred=pred, target=target, mu=mu, log_var=log_var, dpred=dpred, dmu=dmu, dlv=dlv,
                   kld_weight=ctx.kld_weight, input_is_prob=ctx.input_is_prob)
        return dpred, None, dmu, dlv, None, None

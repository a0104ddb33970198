#!/usr/bin/env python
"""Headline benchmark: VAE train samples/sec on synthetic 1x128xT piano-rolls.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # reference CPU arm

A "step" is one pass of the hot path over one batch: forward (models.py:185-188)
+ ELBO (:190-225) + backward (train.py:650) + gradient all-reduce when N > 1 +
the reference's AdamW step over its encoder/decoder groups (train.py:210-228,656).
Workload at every N: BASELINE.json configs[1] -- default widths [32,64,128,256],
L = 10, batch 512 PER GPU (weak scaling), bf16 activations / fp32 master weights.

Prints ONE JSON line (rank 0).  See README / DESIGN.md for the field meanings.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "vae_train_samples_per_sec"
UNIT = "samples/s"
# BASELINE.json configs -> workloads.  "train" (configs[1]/[2]) is the one `metric` is quoted on and the
# default; "wide_long" is configs[3]; "sample" is configs[4] (decode-only, its own metric name).
WORKLOADS = {
    "train": {"hidden": [32, 64, 128, 256], "latent": 10, "height": 128, "width": 128, "batch": 512,
              "flops_train": 0.9419e9, "flops_decode": 0.1985e9},
    "wide_long": {"hidden": [64, 128, 256, 512], "latent": 256, "height": 128, "width": 1024, "batch": 64,
                  "flops_train": 30.728e9, "flops_decode": None},
    "sample": {"hidden": [32, 64, 128, 256], "latent": 10, "height": 128, "width": 128, "batch": 4096,
               "flops_train": 0.9419e9, "flops_decode": 0.1985e9},
}
HIDDEN = WORKLOADS["train"]["hidden"]
LATENT = WORKLOADS["train"]["latent"]
FLOPS_PER_SAMPLE_TRAIN = {}  # filled by apply_workload (FlopCounterMode on the reference)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="train", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="batch per GPU (default: the workload's)")
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true", help="disable CUDA-graph replay of the step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-zero-arena", action="store_true", help="let the library clear every statistics / split-K buffer with its own memset")
    ap.add_argument("--bucket-cap-mb", type=float, default=25.0, help="gradient bucket size (parallel.GradientBuckets)")
    ap.add_argument("--no-grad-sink", action="store_true", help="N = 1: let autograd allocate / accumulate the gradients instead of the flat bucket sink")
    ap.add_argument("--sample-chunk", type=int, default=512, help="sample workload, e2e leg: images per decode / copy chunk")
    ap.add_argument("--torch-adamw", action="store_true", help="step with torch.optim.AdamW instead of the fused kernel")
    ap.add_argument("--cpu-batch", type=int, default=64)
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the torch-eager (reference class on cuda) leg")
    ap.add_argument("--reserve-sms", type=int, default=0, help="SMs left to NCCL kernels in multi-GPU runs (persistent kernels use the rest)")
    ap.add_argument("--profile-json", default=None, help="write the per-kernel CUDA-event breakdown here")
    return apply_workload(ap.parse_args())


def apply_workload(a):
    global HIDDEN, LATENT
    w = WORKLOADS[a.workload]
    HIDDEN, LATENT = list(w["hidden"]), w["latent"]
    a.batch = a.batch or w["batch"]
    a.height = a.height or w["height"]
    a.width = a.width or w["width"]
    if (a.height, a.width) == (w["height"], w["width"]):
        FLOPS_PER_SAMPLE_TRAIN[(a.height, a.width)] = w["flops_train"]
    return a


def activation_bytes(H, W, hidden, elem=2):
    """Bytes per sample of every layer output the step keeps (bf16 NHWC), forward + as many gradients."""
    n, h, w = 0, H, W
    for c in hidden:
        h, w = h // 2, w // 2
        n += c * h * w
    for c in list(reversed(hidden))[1:] + [hidden[0]]:
        h, w = h * 2, w * 2
        n += c * h * w
    return 2 * (n * 2 + H * W) * elem


def workload_name(a):
    return f"VanillaVAE widths {HIDDEN} L={LATENT}, synthetic 1x{a.height}x{a.width} piano-rolls, batch {a.batch}/GPU"


# ----------------------------------------------------------------------------
# clocks: sampled DURING the timed region (B200_PROFILING.md "clocks line")
# ----------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        """Summary of the samples taken while the timed region [t0, t1] ran (nvidia-smi is started before
        the warm-up so that it is already streaming; if the region is shorter than one sampling period the
        samples within 150 ms of it are used)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        picked = self.samples
        if t0 is not None and t1 is not None:
            picked = [x for x in self.samples if t0 <= x[0] <= t1 + 0.02]
            if not picked:
                picked = [x for x in self.samples if t0 - 0.15 <= x[0] <= t1 + 0.15]
        for _, s in picked:
            f = [t.strip() for t in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------
# algorithmic bytes / flops of one C-ABI launch (DESIGN.md "kernels" table)
# ----------------------------------------------------------------------------
def _sz(code):
    return 2 if code == 1 else 4


def algorithmic_cost(name, tag):
    """-> (bytes, flops) one launch must move / compute, from its integer args."""
    t = list(tag)
    if name == "vae_conv_fprop":
        B, Hi, Wi, Ci, Co, s, di, do = t[:8]
        Ho, Wo = Hi // s, Wi // s
        return (B * Hi * Wi * Ci * _sz(di) + B * Ho * Wo * Co * _sz(do) + 9 * Ci * Co * _sz(do), 2 * B * Ho * Wo * Co * 9 * Ci)
    if name == "vae_conv_dgrad":
        B, Hi, Wi, Ci, Co, s, d = t[:7]
        Ho, Wo = Hi // s, Wi // s
        return (B * Hi * Wi * Ci * _sz(d) + B * Ho * Wo * Co * _sz(d) + 9 * Ci * Co * _sz(d), 2 * B * Ho * Wo * Co * 9 * Ci)
    if name == "vae_conv_wgrad":
        B, Hi, Wi, Ci, Co, s, dx, dy = t[:8]
        Ho, Wo = Hi // s, Wi // s
        return (B * Hi * Wi * Ci * _sz(dx) + B * Ho * Wo * Co * _sz(dy) + 9 * Ci * Co * 4, 2 * B * Ho * Wo * Co * 9 * Ci)
    if name in ("vae_conv_umma", "vae_conv_umma_tma", "vae_conv_umma_rw", "vae_conv_umma_rw_act", "vae_conv_umma_tma_act"):
        B, Hi, Wi, Ci, Co, s = t[:6]
        Ho, Wo = Hi // s, Wi // s
        return (2 * (B * Hi * Wi * Ci + B * Ho * Wo * Co + 9 * Ci * Co), 2 * B * Ho * Wo * Co * 9 * Ci)
    if name in ("vae_conv_umma_rw_bnbwd", "vae_conv_umma_tma_bnbwd"):   # + one read of y (shape of the output)
        B, Hi, Wi, Ci, Co, s, mode = t[:7]
        Ho, Wo = Hi // s, Wi // s
        out_elems = B * Ho * Wo * Co if mode == 0 else B * Hi * Wi * Ci
        return (2 * (B * Hi * Wi * Ci + B * Ho * Wo * Co + 9 * Ci * Co + out_elems), 2 * B * Ho * Wo * Co * 9 * Ci)
    if name in ("vae_conv_wgrad_umma", "vae_conv_wgrad_tma"):
        B, Hi, Wi, Ci, Co, s = t[:6]
        Ho, Wo = Hi // s, Wi // s
        return (2 * (B * Hi * Wi * Ci + B * Ho * Wo * Co) + 4 * 9 * Ci * Co, 2 * B * Ho * Wo * Co * 9 * Ci)
    if name == "vae_final_conv_fwd_tc":
        B, H, W, C = t[:4]
        return (B * H * W * (C * 2 + 2 + 4), 2 * B * H * W * 9 * C)
    if name == "vae_final_bwd_reduce_tc":
        B, H, W, C = t[:4]
        return (B * H * W * (C + 1) * 2, 4 * B * H * W * 9 * C)
    if name == "vae_final_bwd_apply_tc":
        B, H, W, C = t[:4]
        return (B * H * W * (2 * C + 1) * 2, 2 * B * H * W * 9 * C)
    if name == "vae_final_conv_fwd":  # reads y, writes logits (+ fp32 probabilities)
        B, H, W, C, d = t[:5]
        return (B * H * W * (C * _sz(d) + _sz(d) + 4), 2 * B * H * W * 9 * C)
    if name == "vae_final_bwd_reduce":  # reads y and dlogits
        B, H, W, C, d = t[:5]
        return (B * H * W * (C + 1) * _sz(d), 4 * B * H * W * 9 * C)
    if name == "vae_final_bwd_apply":  # reads y and dlogits, writes dy
        B, H, W, C, _, d = t[:6]
        return (B * H * W * (2 * C + 1) * _sz(d), 2 * B * H * W * 9 * C)
    if name == "vae_bn_finalize_apply_lrelu":  # integer args: n_pix, C, count, training, dtype
        n, C = t[:2]
        d = t[4] if len(t) > 4 else 1
        return (2 * n * C * _sz(d), 3 * n * C)
    if name in ("vae_skinny_reduce", "vae_skinny_wgrad", "vae_skinny_expand"):
        M, N, K = t[:3]
        return (2 * M * K + 4 * N * K + 4 * M * N, 2 * M * N * K)
    if name == "vae_skinny_reduce_pair":   # one read of A, two weight matrices, two outputs
        M, N, K = t[:3]
        return (2 * M * K + 8 * N * K + 8 * M * N, 4 * M * N * K)
    if name == "vae_bn_apply_lrelu":
        n, C, d = t[:3]
        return (2 * n * C * _sz(d), 3 * n * C)
    if name == "vae_bn_lrelu_bwd_reduce":
        n, C, d = t[:3]
        return (2 * n * C * _sz(d), 8 * n * C)
    if name == "vae_bn_lrelu_bwd_fused":   # n_pix, C, training, dtype: d_act + y in (the second read is served by the L2), dy out
        n, C = t[:2]
        d = t[3] if len(t) > 3 else 0
        return (3 * n * C * _sz(d), 18 * n * C)
    if name == "vae_bn_lrelu_bwd_apply":
        n, C = t[:2]
        d = t[3] if len(t) > 3 else 0
        return (3 * n * C * _sz(d), 10 * n * C)
    if name in ("vae_elbo_fwd", "vae_elbo_bwd"):
        n, B, L, d = t[0], t[1], t[2], t[3]
        per = (_sz(d) + 4) if name == "vae_elbo_fwd" else (2 * _sz(d) + 4)
        return (n * per + 6 * B * L * 4, 20 * n)
    if name == "vae_elbo_fused":  # logits + fp32 target in, dlogits out, six latent vectors
        n, B, L, d = t[0], t[1], t[2], t[3]
        return (n * (2 * _sz(d) + 4) + 6 * B * L * 4, 30 * n)
    if name == "vae_sigmoid":
        n, d = t[:2]
        return (n * (_sz(d) + 4), 4 * n)
    if name in ("vae_nhwc_to_nchw", "vae_nchw_to_nhwc"):
        B, HW, C, ds, dd = t[:5]
        return (B * HW * C * (_sz(ds) + _sz(dd)), 0)
    if name in ("vae_linear_reduce_tc", "vae_linear_wgrad_tc"):   # M, L, F: fp32 weight / gradient + bf16 activations
        M, L, F = t[:3]
        return (4 * L * F + 2 * M * F + 4 * M * L, 2 * M * L * F)
    if name == "vae_linear_expand_tc":   # M, L, F, layout, dtype of the output
        M, L, F = t[:3]
        return (4 * L * F + _sz(t[4] if len(t) > 4 else 1) * M * F + 4 * M * L, 2 * M * L * F)
    if name == "vae_gemm":
        M, N, K = t[:3]
        return (4 * (M * K + K * N + M * N), 2 * M * N * K)
    return (0, 0)


# DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels that can dominate
# the B=512 step, from ONE `ncu --set full` capture per kernel (scripts/profile_top.py): the output-block
# kernels from profiles/r04k_ncu_full_output_block.csv (final code of round 2), the conv kernels from
# profiles/r01e_ncu_full_top_kernels.csv.
# Keyed by (C-ABI entry point, leading integer arguments).
NCU_TRAFFIC_BYTES = {
    ("vae_final_conv_fwd_tc", (512, 128, 128, 32)): 559.68e6 + 45.58e6,     # profiles/r04k_ncu_full_output_block.csv
    ("vae_final_bwd_reduce_tc", (512, 128, 128, 32)): 553.80e6 + 4.50e6,
    ("vae_final_bwd_apply_tc", (512, 128, 128, 32)): 556.37e6 + 499.85e6,
    ("vae_conv_umma_rw", (512, 128, 128, 32, 32, 2, 1)): 134.33e6 + 477.68e6,
    ("vae_conv_umma_rw", (512, 128, 128, 32, 32, 2, 0)): 536.98e6 + 118.08e6,
    ("vae_conv_wgrad_tma", (512, 128, 128, 32, 32, 2)): 671.55e6 + 6.52e6,
}


def measured_traffic(name, tag):
    for (n, prefix), b in NCU_TRAFFIC_BYTES.items():
        if n == name and tuple(tag[:len(prefix)]) == prefix:
            return b
    return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ----------------------------------------------------------------------------
# per-kernel CUDA-event breakdown and the roofline object of the dominant kernel
# ----------------------------------------------------------------------------
def kernel_breakdown(ops, step_fn, profile_json=None, n_prof=3):
    """Every C-ABI launch of `n_prof` eager steps bracketed with CUDA events on the launching stream.
    -> (rows sorted by time [(key, launches, ms)], kernel ms per step, launches per step, n_prof)."""
    for _ in range(2):
        step_fn()
    torch.cuda.synchronize()
    tl = ops.Timeline()
    ops.set_timeline(tl)
    l0 = ops.LAUNCHES
    for _ in range(n_prof):
        step_fn()
    torch.cuda.synchronize()
    launches_per_step = (ops.LAUNCHES - l0) // n_prof
    ops.set_timeline(None)
    summ = tl.summary()
    rows = sorted(((k, n, ms) for k, (n, ms) in summ.items()), key=lambda r: -r[2])
    total_kernel_ms = sum(r[2] for r in rows) / n_prof
    if profile_json:
        with open(profile_json, "w") as f:
            json.dump({"per_step_kernel_ms": total_kernel_ms, "launches_per_step": launches_per_step,
                       "kernels": [{"name": k[0], "args": list(k[1]), "launches_per_step": n / n_prof,
                                    "ms_per_step": ms / n_prof, "share": ms / n_prof / total_kernel_ms,
                                    "alg_bytes": algorithmic_cost(*k)[0], "alg_flops": algorithmic_cost(*k)[1]} for k, n, ms in rows]}, f, indent=1)
    return rows, total_kernel_ms, launches_per_step, n_prof


def roofline_of(ops, top_key, step_fn, reps, share):
    tl2 = ops.Timeline(only=top_key)
    ops.set_timeline(tl2)
    for _ in range(reps):
        step_fn()
    torch.cuda.synchronize()
    ops.set_timeline(None)
    n_top, ms_top = tl2.summary()[top_key]
    top_bytes, top_flops = algorithmic_cost(*top_key)
    peaks = load_peaks()
    avg_s = ms_top / n_top / 1e3
    hbm_bound = top_bytes / (peaks["hbm_gbs"] * 1e9) >= top_flops / (peaks["bf16_tflops_sustained"] * 1e12)
    if hbm_bound:
        achieved, peak, unit = top_bytes / avg_s / 1e9, peaks["hbm_gbs"], "GB/s"
    else:
        achieved, peak, unit = top_flops / avg_s / 1e12, peaks["bf16_tflops_sustained"], "TFLOP/s"
    return {"bound": "hbm" if hbm_bound else "tensor", "achieved": achieved, "peak": peak, "unit": unit,
            "frac": achieved / peak, "traffic": measured_traffic(*top_key), "kernel": top_key[0], "kernel_args": list(top_key[1]),
            "avg_launch_us": avg_s * 1e6, "alg_bytes_per_launch": top_bytes, "alg_flops_per_launch": top_flops,
            "share_of_step_kernel_time": share,
            "peak_source": peaks["source"] + (" (sustained bf16 figure: the kernel is timed inside a long step)" if not hbm_bound else "")}


# ----------------------------------------------------------------------------
# CPU arms
# ----------------------------------------------------------------------------
def cpu_reference_run(a, steps, warmup, batch):
    """The reference's CPU train step (train.py:635-656) on all host cores: the UNMODIFIED reference class
    from baseline/_ref (kind "reference") when it is installed, else the oracle port (kind "port")."""
    from oracle import reference_driver as rd
    from oracle import vae_oracle as vo

    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    x, _ = vo.synthetic_batch(batch, a.height, a.width, LATENT)
    if rd.available():
        trainer = rd.ReferenceTrainStep(a.height, a.width, HIDDEN, LATENT, batch, device="cpu")
        kind = "reference"
        what = ("midi_autoencoder.models.VanillaVAE from baseline/_ref driven as train.py:635-656"
                + ("" if (a.height, a.width) == (32, 32) else " (literals models.py:33,166 generalised, SURVEY 8c)"))
    else:
        params = vo.xavier_params(1, LATENT, a.height, a.width, HIDDEN, seed=0)
        trainer = vo.CpuTrainStep(params, batch)
        kind = "port"
        what = "oracle/vae_oracle.CpuTrainStep = reference train.py:635-656 with ATen CPU ops"
    for _ in range(warmup):
        trainer(x)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        trainer(x)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": batch * steps / total, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{steps} steps of batch {batch} fp32 1x{a.height}x{a.width} (median {statistics.median(times)*1e3:.0f} ms/step), {what}",
            "ms_per_step": 1e3 * total / steps}


def cpu_sample_run(a, steps, warmup, batch):
    """Reference `sample()` (models.py:250-263) on the host cores, eval mode, no_grad."""
    from oracle import reference_driver as rd
    from oracle import vae_oracle as vo

    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    if rd.available():
        torch.manual_seed(0)
        model = rd.build_model(a.height, a.width, HIDDEN, LATENT).eval()
        fn = lambda: model.sample(batch, "cpu")
        kind, what = "reference", "midi_autoencoder.models.VanillaVAE.sample from baseline/_ref (eval, no_grad)"
    else:
        params = vo.xavier_params(1, LATENT, a.height, a.width, HIDDEN, seed=0)
        fn = lambda: vo.decode_only(params, torch.randn(batch, LATENT), a.height, a.width)
        kind, what = "port", "oracle/vae_oracle.decode_only"
    times = []
    with torch.no_grad():
        for _ in range(warmup):
            fn()
        for _ in range(steps):
            t0 = time.perf_counter()
            fn()
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": batch * steps / total, "unit": "images/s", "cores": cores, "kind": kind,
            "sample": f"{steps} calls of sample({batch}) fp32 1x{a.height}x{a.width} (median {statistics.median(times)*1e3:.0f} ms), {what}",
            "ms_per_step": 1e3 * total / steps}


def gpu_eager_run(a, dev, batch, steps=10, warmup=3):
    """SURVEY.md section 2b / BASELINE configs[1] "vs reference torch eager": the UNMODIFIED reference module
    on the SAME GPU through cuDNN / cuBLAS, fp32 (TF32 convolutions: torch's default) and bf16 autocast, timed
    with CUDA events.  A second stated baseline beside cpu_baseline; None when baseline/_ref is absent."""
    from oracle import reference_driver as rd
    from oracle import vae_oracle as vo

    if not rd.available():
        return None
    out = {"kind": "reference", "what": "midi_autoencoder.models.VanillaVAE from baseline/_ref on cuda (ATen eager: cuDNN/cuBLAS), "
                                        "train.py:635-656 with torch.optim.AdamW over encoder+decoder", "batch": batch, "steps": steps}
    x = vo.synthetic_batch(batch, a.height, a.width, LATENT)[0].to(dev)
    for name, autocast in (("fp32_tf32conv", None), ("bf16_autocast", torch.bfloat16)):
        try:
            trainer = rd.ReferenceTrainStep(a.height, a.width, HIDDEN, LATENT, batch, device=dev, autocast=autocast)
            for _ in range(warmup):
                trainer(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                trainer(x)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"value": batch / (ms / 1e3), "unit": UNIT, "ms_per_step": ms}
            del trainer
        except Exception as e:  # report, do not hide (e.g. out of memory at the wide config)
            out[name] = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
        torch.cuda.empty_cache()
    return out


def gpu_eager_sample_run(a, dev, batch, steps=10, warmup=3):
    from oracle import reference_driver as rd

    if not rd.available():
        return None
    torch.manual_seed(0)
    model = rd.build_model(a.height, a.width, HIDDEN, LATENT).to(dev).eval()
    out = {"kind": "reference", "what": "midi_autoencoder.models.VanillaVAE.sample from baseline/_ref on cuda (eval, no_grad; "
                                        "z drawn on the host as models.py:258 does)", "batch": batch, "steps": steps}
    for name, autocast in (("fp32_tf32conv", None), ("bf16_autocast", torch.bfloat16)):
        try:
            with torch.no_grad(), torch.autocast("cuda", dtype=autocast or torch.float32, enabled=autocast is not None):
                for _ in range(warmup):
                    model.sample(batch, dev)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    model.sample(batch, dev)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"value": batch / (ms / 1e3), "unit": "images/s", "ms_per_step": ms}
        except Exception as e:
            out[name] = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
        torch.cuda.empty_cache()
    return out


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(a.steps, 30)), max(1, min(a.warmup, 5))
    if a.workload == "sample":
        r = cpu_sample_run(a, max(1, min(a.steps, 10)), 1, min(a.batch, 256))
        metric, unit = "vae_sample_images_per_sec", "images/s"
        note = f"CPU arm runs a bounded sample: sample({min(a.batch, 256)}) per step on {r['cores']} host cores"
    else:
        cpu_batch = a.cpu_batch if a.workload == "train" else min(a.cpu_batch, 8)
        if a.workload != "train":
            steps, warmup = min(steps, 5), 1
        r = cpu_reference_run(a, steps, warmup, cpu_batch)
        metric, unit = METRIC, UNIT
        note = f"CPU arm runs a bounded sample: batch {cpu_batch} per step on {r['cores']} host cores"
    line = {
        "impl": "reference", "metric": metric, "value": r["value"], "unit": unit, "n_gpus": a.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "note": note},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------
def b200_arm(a):
    import torch.distributed as dist

    import torch_vae_b200 as tv
    from torch_vae_b200 import ops
    from torch_vae_b200.parallel import GradientBuckets
    from oracle import vae_oracle as vo  # synthetic-input generator + CPU baseline leg only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; torch_vae_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    tv.load_library()
    if world > 1 and a.reserve_sms > 0:
        ops.set_reserved_sms(a.reserve_sms)

    dtype = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    B, H, W = a.batch, a.height, a.width
    torch.manual_seed(0)
    model = tv.VanillaVAE(1, LATENT, (H, W), hidden_dims=list(HIDDEN), compute_dtype=dtype).to(dev)
    model.manual_seed(0, rank)
    model.train()
    lr = 0.01 * (B * world) / 128  # train.py:201
    groups = [{"params": model.encoder.parameters(), "lr": lr, "name": "encoder"},
              {"params": model.decoder.parameters(), "lr": lr, "name": "decoder"}]  # train.py:210-228
    if a.torch_adamw:
        opt = torch.optim.AdamW(groups, lr=lr, weight_decay=0.0, capturable=not a.no_graph)
        sched = None
    else:
        opt = tv.FusedAdamW(groups, lr=lr, weight_decay=0.0)
        # train.py:232-238: OneCycleLR over the groups' lrs, stepped per batch (:659); here advanced on the
        # device so that the captured step follows it.  1000 epochs x 100 steps: the schedule never runs out.
        sched = tv.optim.FusedOneCycleLR(opt, [g["lr"] for g in opt.param_groups], epochs=1000, steps_per_epoch=100)
    # gradient sink: the backward kernels write every gradient straight into flat buckets (no per-parameter
    # AccumulateGrad launches); at N > 1 the buckets are all-reduced (NCCL AVG) under the rest of backward, at N = 1 the
    # sink is local.  --no-grad-sink at N = 1: plain autograd gradient tensors.
    buckets = None
    if world > 1 or not a.no_grad_sink:
        buckets = GradientBuckets(model, bucket_cap_mb=a.bucket_cap_mb, first_bucket_cap_mb=min(1.0, a.bucket_cap_mb))
        buckets.broadcast_parameters(0)

    x_host, _ = vo.synthetic_batch(B, H, W, LATENT, seed=1234 + rank)
    x_host = x_host.pin_memory()
    x_dev = x_host.to(dev)
    losses_host = torch.empty(3, dtype=torch.float32).pin_memory()

    if not a.no_zero_arena:
        ops.enable_zero_arena(dev)   # statistics / split-K scratch from one arena, cleared once per step

    def train_step(x):
        ops.zero_arena_reset(dev)
        out = model.forward(x)  # train.py:635
        opt.zero_grad(set_to_none=True)  # :642
        losses = model.loss(out)  # :644
        if buckets is not None:
            buckets.begin()
        losses["loss"].backward()  # :650
        if buckets is not None:
            buckets.finish()
        opt.step()  # :656
        if sched is not None:
            sched.step()  # :659
        return losses

    # -- per-kernel breakdown (eager, untimed): find the dominant kernel -----
    rows, total_kernel_ms, launches_per_step, n_prof = kernel_breakdown(ops, lambda: train_step(x_dev), a.profile_json if rank == 0 else None)
    top_key = rows[0][0]

    # -- optional CUDA-graph capture of the whole step -------------------------
    use_graph = not a.no_graph  # NCCL all-reduces on the side stream are captured with the step
    graph = None
    static_x = x_dev.clone()
    if use_graph:
        try:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(3):
                    train_step(static_x)
            torch.cuda.current_stream().wait_stream(s)
            graph = torch.cuda.CUDAGraph()
            opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(graph):
                static_losses = train_step(static_x)
            torch.cuda.synchronize()
        except Exception as e:  # report, do not hide
            if rank == 0:
                print(f"[bench] CUDA-graph capture failed, running eagerly: {e}", file=sys.stderr)
            graph = None
            use_graph = False

    def run_step(x):
        if graph is not None:
            if x is not static_x:
                static_x.copy_(x, non_blocking=True)
            graph.replay()
            return static_losses
        return train_step(x)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # -- device-resident timing (value) -----------------------------------------
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(max(3, a.warmup)):
        run_step(x_dev)
    if rank == 0:  # let nvidia-smi reach its streaming state before the (possibly very short) timed region
        t_wait = time.time()
        while not clocks.samples and time.time() - t_wait < 2.0:
            time.sleep(0.01)  # (no extra steps here: with N > 1 every step is a collective)
    l0 = ops.LAUNCHES
    t_region0 = time.time()
    ms_total = timed(lambda: run_step(x_dev), a.steps)
    t_region1 = time.time()
    eager_launches = ops.LAUNCHES - l0
    clk = clocks.stop(t_region0, t_region1) if rank == 0 else None
    ms_per_step = ms_total / a.steps
    value = B * world * a.steps / (ms_total / 1e3)

    # -- dominant kernel, bracketed live (eager so the events land on the launch stream)
    roofline = roofline_of(ops, top_key, lambda: train_step(x_dev), max(3, min(a.steps, 10)),
                           rows[0][2] / n_prof / total_kernel_ms)

    # -- end to end through the public API with HOST buffers (e2e) ---------------
    # What a training loop around the drop-in does per batch (train.py:620-674): the batch sits in
    # pinned HOST memory as uint8 piano-rolls (what a DataLoader hands over), is copied to the device,
    # goes through the device-side transform of torch_vae_b200.data_transformations (north_star item 3:
    # crop / scale / normalise on the GPU instead of in DataLoader workers), the train step runs, and
    # the loss is read back on the host every step (train.py:672).  The copy of batch i+1 is issued on
    # a side stream before step i runs (double buffering, as a prefetching loader does); every timed
    # step still contains one full H2D copy and one D2H read.
    from torch_vae_b200.data_transformations import get_transform

    rolls_host = (x_host[:, 0] * 255.0).to(torch.uint8).contiguous().pin_memory()
    train_tf, _ = get_transform("midi", (H, W), {"normalization": "piano-roll", "seed": 1234 + rank})
    dev_rolls = [torch.empty(rolls_host.shape, dtype=torch.uint8, device=dev) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    e2e_state = {"i": 0}
    e2e_loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    loss_done = [torch.cuda.Event() for _ in range(2)]

    def prefetch(k):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[k])  # the transform that last read this buffer has finished
            dev_rolls[k].copy_(rolls_host, non_blocking=True)
            ready[k].record(copy_stream)

    prefetch(0)

    def e2e_step():
        k = e2e_state["i"] & 1
        e2e_state["i"] += 1
        prefetch(k ^ 1)
        cur = torch.cuda.current_stream()
        cur.wait_event(ready[k])
        # the transform writes straight into the step's input buffer (the captured graph's static input)
        xd = train_tf(dev_rolls[k], out=static_x if graph is not None else None)
        freed[k].record(cur)
        losses = run_step(xd)
        e2e_loss_host[k:k + 1].copy_(losses["loss"].detach().reshape(1), non_blocking=True)
        loss_done[k].record(cur)
        # The caller reads a loss every step (train.py:672), one step behind: step i's launches are queued
        # before step i-1's loss is awaited, so the GPU never idles through a host round trip.
        if e2e_state["i"] == 1:
            return None
        loss_done[k ^ 1].synchronize()
        return float(e2e_loss_host[k ^ 1])

    for _ in range(3):
        e2e_step()
    ms_e2e = timed(e2e_step, a.steps)
    e2e_value = B * world * a.steps / (ms_e2e / 1e3)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": a.dtype if a.dtype != "fp32" else "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "height": H, "width": W, "batch_per_gpu": B, "global_batch": B * world,
                       "parallelism": f"dp{world}", "cuda_graph": bool(graph is not None),
                       "optimizer": ("torch.optim.AdamW" if a.torch_adamw else "FusedAdamW + FusedOneCycleLR (device-resident lr/beta1, torch.optim.AdamW state layout)")
                                    + " over encoder+decoder groups (train.py:210-238), optimizer.step and scheduler.step inside the timed step",
                       "gradients": (("backward kernels write straight into flat buckets (parallel.GradientBuckets: "
                                      + ("NCCL AVG all-reduce per bucket under the rest of backward)" if world > 1 else "local sink, no collective)"))
                                     if buckets is not None else "autograd gradient tensors"),
                       "l2": f"per-step activation working set {activation_bytes(H, W, HIDDEN) * B / 1e6:.0f} MB >> 126 MB L2 (inputs larger than L2)"},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": rolls_host.numel() * world, "d2h_bytes_per_step": 4 * world,
                    "ms_per_step": ms_e2e / a.steps,
                    "input": "uint8 rolls in pinned host memory -> H2D (prefetched one batch ahead) -> device RollTransform -> step -> loss D2H (read every step, one step behind)"},
            "gpu_launches": launches_per_step * a.steps,
            "gpu_launches_per_step": launches_per_step,
            "roofline": roofline,
            "step_tflops": FLOPS_PER_SAMPLE_TRAIN.get((H, W), 0) * value / 1e12 if (H, W) in FLOPS_PER_SAMPLE_TRAIN else None,
        }
        if world == 1 and not a.no_eager_baseline:
            del graph
            torch.cuda.empty_cache()
            eager = gpu_eager_run(a, dev, B)
            if eager is not None:
                line["gpu_eager_baseline"] = eager
        if world == 1 and not a.no_cpu_baseline:
            if a.workload == "train":
                r = cpu_reference_run(a, 10, 2, a.cpu_batch)
            else:
                r = cpu_reference_run(a, 3, 1, min(a.cpu_batch, 8))
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Captured graphs hold NCCL kernels; tearing the communicator down under them can block, so
        # the ranks meet at a barrier, drain the device and leave without running destructors.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def b200_sample_arm(a):
    """BASELINE.json configs[4]: decoder-only sampling, z ~ N(0, I) -> piano-rolls (models.py:250-263),
    batch 4096 on one B200, eval mode, no_grad.  `value`: z from the device Philox stream, images stay in
    HBM.  `e2e`: the reference's own call shape -- z drawn on the HOST with torch.randn (models.py:258),
    copied to the device, decoded, and the images copied back to pinned host memory."""
    import torch_vae_b200 as tv
    from torch_vae_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; torch_vae_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    tv.load_library()
    dtype = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    B, H, W = a.batch, a.height, a.width
    torch.manual_seed(0)
    model = tv.VanillaVAE(1, LATENT, (H, W), hidden_dims=list(HIDDEN), compute_dtype=dtype).to(dev).eval()
    model.manual_seed(0, rank)

    def sample_step():
        with torch.no_grad():
            return model.sample(B, dev, device_rng=True)

    rows, total_kernel_ms, launches_per_step, n_prof = kernel_breakdown(ops, sample_step, a.profile_json if rank == 0 else None)
    top_key = rows[0][0]

    graph, static_out = None, None
    if not a.no_graph:
        try:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(3):
                    sample_step()
            torch.cuda.current_stream().wait_stream(s)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = sample_step()
            torch.cuda.synchronize()
        except Exception as e:
            if rank == 0:
                print(f"[bench] CUDA-graph capture failed, running eagerly: {e}", file=sys.stderr)
            graph = None

    def run_step():
        if graph is not None:
            graph.replay()
            return static_out
        return sample_step()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(max(3, a.warmup)):
        run_step()
    if rank == 0:
        t_wait = time.time()
        while not clocks.samples and time.time() - t_wait < 2.0:
            time.sleep(0.01)
    t0 = time.time()
    ms_total = timed(run_step, a.steps)
    t1 = time.time()
    clk = clocks.stop(t0, t1) if rank == 0 else None
    value = B * world * a.steps / (ms_total / 1e3)
    roofline = roofline_of(ops, top_key, sample_step, max(3, min(a.steps, 10)), rows[0][2] / n_prof / total_kernel_ms)

    # e2e: host z -> H2D -> decode -> D2H of the images
    img_host = torch.empty((B, 1, H, W), dtype=torch.float32).pin_memory()
    z_host = torch.empty((B, LATENT), dtype=torch.float32).pin_memory()

    def e2e_step():
        torch.randn((B, LATENT), out=z_host)  # models.py:258
        # decode in chunks, each chunk's D2H copy under the next chunk's kernels; returns with the images on the host
        model.sample_to_host(B, out=img_host, z=z_host, chunk=a.sample_chunk)

    for _ in range(3):
        e2e_step()
    ms_e2e = timed(e2e_step, a.steps)
    if rank == 0:
        line = {
            "metric": "vae_sample_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(3, a.warmup), "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": a.dtype if a.dtype != "fp32" else "f32", "data": "synthetic",
            "config": {"workload": "decoder-only sampling (BASELINE configs[4]): " + workload_name(a), "height": H, "width": W,
                       "batch_per_gpu": B, "parallelism": f"replicas x{world}", "cuda_graph": graph is not None,
                       "l2": f"{B * H * W * 32 * 2 / 1e6:.0f} MB full-resolution activation per step >> 126 MB L2"},
            "clocks": clk,
            "e2e": {"value": B * world * a.steps / (ms_e2e / 1e3), "unit": "images/s", "h2d_bytes_per_step": z_host.numel() * 4 * world,
                    "d2h_bytes_per_step": img_host.numel() * 4 * world, "ms_per_step": ms_e2e / a.steps,
                    "input": f"z = torch.randn on the host (models.py:258) -> H2D -> VanillaVAE.sample_to_host: decode in chunks of {a.sample_chunk}, "
                             "fp32 images D2H into pinned memory on a copy stream under the next chunk's decode"},
            "gpu_launches": launches_per_step * a.steps, "gpu_launches_per_step": launches_per_step,
            "roofline": roofline,
            "step_tflops": WORKLOADS[a.workload]["flops_decode"] * value / 1e12 if (H, W) == (128, 128) else None,
        }
        if world == 1 and not a.no_eager_baseline:
            del graph
            torch.cuda.empty_cache()
            eager = gpu_eager_sample_run(a, dev, B)
            if eager is not None:
                line["gpu_eager_baseline"] = eager
        if world == 1 and not a.no_cpu_baseline:
            r = cpu_sample_run(a, 3, 1, min(B, 256))
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


def main():
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    elif a.workload == "sample":
        b200_sample_arm(a)
    else:
        b200_arm(a)


if __name__ == "__main__":
    main()

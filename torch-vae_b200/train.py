"""Launcher for the B200-native VAE training path.

    python -m torch_vae_b200.train --config configs/vae.yaml --epochs 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
           -m torch_vae_b200.train --batch-size 512 --image-size 128 --dtype bf16

replaces ``scripts/mnist.sh:120-160`` + the subset of ``train.py:770-1063`` flags that concerns the hot path:
the model (``--n_features``, ``--image-size``), the optimisation routine of train.py:195-238 (``--epochs``,
``--lr`` per 128 samples scaled by the global batch, ``--lr-encoder-mult``, ``--lr-decoder-mult``,
``--weight-decay``, AdamW over the encoder / decoder groups, OneCycleLR stepped per batch), ``--batch-size`` per
GPU, ``--seed``, ``--checkpoint`` / ``--models-dir`` (the ``safe_save_model`` file layout of utils.py:311-351
with the contents of train.py:445-460), ``--log-interval``; and the ``configs/vae.yaml`` schema
(``model_params`` / ``exp_params`` / ``data_params`` / ``trainer_params`` / ``logging_params``) through
``--config``: YAML values are the defaults, explicit flags win.

One process per GPU (RANK / LOCAL_RANK / WORLD_SIZE from torchrun), NCCL gradient buckets overlapped with
backward, the whole step replayed as a CUDA graph, piano-rolls staged once in HBM (``--data`` a ``.npy`` /
``.pt`` file of uint8 ``[N, 128, T]`` rolls, or ``synthetic``).  No CPU fallback: without a CUDA device it raises.
Datasets other than piano-rolls, wandb logging and the evaluation metrics of evaluation.py stay with the
reference's own driver (INTEGRATION.md shows the two-line change that makes it use this package's model).
"""

from __future__ import annotations

import argparse
import os
import sys
import time
from typing import Optional

import torch

BASE_BATCH_SIZE = 128   # train.py:17


def get_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="torch_vae_b200.train", description="Train the VanillaVAE on piano-rolls (B200 path).")
    p.add_argument("--config", type=str, default=None, help="configs/vae.yaml-style file; flags given explicitly win")
    g = p.add_argument_group("Dataset")
    g.add_argument("--data", type=str, default="synthetic", help=".npy / .pt file of uint8 rolls [N, pitches, steps], or 'synthetic'")
    g.add_argument("--n-rolls", type=int, default=8192, help="synthetic rolls to generate")
    g.add_argument("--transform-type", type=str, default="midi", help="midi | noaug | digits (data_transformations.get_transform)")
    g.add_argument("--image-size", type=str, default="128", help="S or HxW (pitches x time steps); default 128")
    g.add_argument("--max-pitch-shift", type=int, default=0)
    g.add_argument("--binarize-threshold", type=float, default=0.5)
    g = p.add_argument_group("Architecture")
    g.add_argument("--model", "--arch", dest="arch", type=str, default="VanillaVAE")
    g.add_argument("--n_features", type=int, default=None, help="latent size (default 10, or model_params.latent_dim)")
    g.add_argument("--hidden-dims", type=str, default=None, help="comma separated widths (default 32,64,128,256)")
    g.add_argument("--kld-weight", type=float, default=None)
    g.add_argument("--dtype", type=str, default="bf16", choices=["bf16", "fp32"], help="activation dtype of the kernels")
    g = p.add_argument_group("Optimization routine")
    g.add_argument("--epochs", type=int, default=None)
    g.add_argument("--lr", dest="lr_relative", type=float, default=None,
                   help=f"maximum learning rate per {BASE_BATCH_SIZE} samples, scaled by the global batch (train.py:201)")
    g.add_argument("--lr-encoder-mult", type=float, default=1.0)
    g.add_argument("--lr-decoder-mult", type=float, default=1.0)
    g.add_argument("--weight-decay", "--wd", dest="weight_decay", type=float, default=None)
    g.add_argument("--optimizer", type=str, default="AdamW")
    g.add_argument("--scheduler", type=str, default="OneCycle")
    g.add_argument("--batch-size", dest="batch_size_per_gpu", type=int, default=None)
    g.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying a CUDA graph")
    g = p.add_argument_group("Output / reproducibility")
    g.add_argument("--models-dir", type=str, default="models")
    g.add_argument("--checkpoint", dest="checkpoint_path", type=str, default="")
    g.add_argument("--seed", type=int, default=None)
    g.add_argument("--log-interval", type=int, default=None)
    g.add_argument("--max-steps", type=int, default=None, help="stop after this many optimisation steps (smoke runs)")
    return p


def _apply_yaml(config: argparse.Namespace) -> None:
    """configs/vae.yaml:1-34 -> defaults of the flags that were not given."""
    from .config import DEFAULTS, load_config

    y = load_config(config.config) if config.config else {k: dict(v) for k, v in DEFAULTS.items()}
    mp, ep, dp, tp, lp = y["model_params"], y["exp_params"], y["data_params"], y["trainer_params"], y["logging_params"]
    from_yaml = config.config is not None
    if config.n_features is None:
        config.n_features = int(mp["latent_dim"]) if from_yaml else 10                      # train.py:877 default 10
    if config.hidden_dims is None:
        config.hidden_dims = list(mp["hidden_dims"])
    else:
        config.hidden_dims = [int(v) for v in str(config.hidden_dims).split(",")]
    if config.kld_weight is None:
        config.kld_weight = float(ep["kld_weight"]) if from_yaml else 1.0                   # models.py:17 default 1.0
    if config.batch_size_per_gpu is None:
        config.batch_size_per_gpu = int(dp["train_batch_size"]) if from_yaml else BASE_BATCH_SIZE
    if config.weight_decay is None:
        config.weight_decay = float(ep["weight_decay"]) if from_yaml else 0.0               # train.py:930 default 0
    if config.seed is None:
        config.seed = int(ep["manual_seed"]) if from_yaml else 0
    if config.epochs is None:
        config.epochs = int(tp.get("max_epochs", 5)) if from_yaml else 5                    # train.py:886 default 5
    if config.log_interval is None:
        config.log_interval = int(lp.get("log_interval", 10))
    # the YAML's learning_rate is absolute; the CLI's --lr is per 128 samples (train.py:201)
    config.lr_absolute = float(ep["learning_rate"]) if (from_yaml and config.lr_relative is None) else None
    if config.lr_relative is None:
        config.lr_relative = 0.01
    if not config.checkpoint_path and from_yaml and lp.get("resume_training") and lp.get("resume_chkpt_path"):
        config.checkpoint_path = str(lp["resume_chkpt_path"])


def _image_size(text: str):
    parts = str(text).lower().replace(",", "x").split("x")
    if len(parts) == 1:
        return int(parts[0]), int(parts[0])
    return int(parts[0]), int(parts[1])


def _load_rolls(config, H: int, W: int, device) -> torch.Tensor:
    from .trainer import synthetic_rolls

    if config.data == "synthetic":
        return synthetic_rolls(config.n_rolls, H, W + 16, device, seed=1234)
    if config.data.endswith(".npy"):
        import numpy as np

        rolls = torch.from_numpy(np.load(config.data))
    else:
        rolls = torch.load(config.data, map_location="cpu")
    if rolls.dim() == 4 and rolls.shape[1] == 1:
        rolls = rolls[:, 0]
    if rolls.dtype != torch.uint8:
        rolls = (rolls.float().clamp(0, 1) * 255).to(torch.uint8) if rolls.is_floating_point() else rolls.to(torch.uint8)
    return rolls.to(device)   # staged ONCE: no per-step H2D afterwards (train.py:625)


def save_checkpoint(path: str, model, optimizer, scheduler, config, **extra) -> None:
    """The file ``utils.safe_save_model`` writes for train.py:445-460 (state_dicts of encoder, decoder,
    optimizer, scheduler + config and counters; temporary file, then an atomic rename).  The reference leaves
    fc_mu / fc_var / decoder_input / final_layer out of its checkpoints; they are added under extra keys so that
    a resume restores the whole model while the reference's own loader (train.py:325-328) still finds its keys."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    head, tail = os.path.split(os.path.abspath(path))
    tmp = os.path.join(head, ".tmp." + tail)
    data = {"encoder": model.encoder.state_dict(), "decoder": model.decoder.state_dict(),
            "optimizer": optimizer.state_dict(), "scheduler": scheduler.state_dict() if scheduler is not None else None,
            "model": model.state_dict(), "config": config}
    data.update(extra)
    torch.save(data, tmp)
    os.rename(tmp, path)


def run(config: argparse.Namespace) -> dict:
    import torch.distributed as dist

    import torch_vae_b200 as tv
    from .data_transformations import get_transform
    from .parallel import GradientBuckets
    from .trainer import DeviceRollDataset, TrainStep

    if config.arch != "VanillaVAE":
        raise NotImplementedError(f"Model {config.arch} not supported.")
    if config.optimizer != "AdamW":
        raise NotImplementedError(f"Optimizer {config.optimizer} not supported (the fused step implements AdamW).")
    if config.scheduler.lower() != "onecycle":
        raise NotImplementedError(f"Scheduler {config.scheduler} not supported.")   # train.py:240
    if not torch.cuda.is_available():
        raise RuntimeError("torch_vae_b200.train needs a CUDA device (there is no CPU fallback)")
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=device)
    tv.load_library()
    _apply_yaml(config)
    H, W = _image_size(config.image_size)
    config.batch_size = config.batch_size_per_gpu * world                                   # train.py:165-166
    config.lr = config.lr_absolute if config.lr_absolute is not None else config.lr_relative * config.batch_size / BASE_BATCH_SIZE
    say = print if rank == 0 else (lambda *a, **k: None)

    torch.manual_seed(config.seed)
    dtype = torch.bfloat16 if config.dtype == "bf16" else torch.float32
    model = tv.VanillaVAE(1, config.n_features, (H, W), hidden_dims=list(config.hidden_dims), kld_weight=config.kld_weight,
                          compute_dtype=dtype).to(device)
    model.manual_seed(config.seed, rank)   # Philox stream of the reparameterisation: seed + rank (train.py:356-363)
    say(model)
    groups = [{"params": model.encoder.parameters(), "lr": config.lr * config.lr_encoder_mult, "name": "encoder"},
              {"params": model.decoder.parameters(), "lr": config.lr * config.lr_decoder_mult, "name": "decoder"}]
    optimizer = tv.FusedAdamW(groups, lr=config.lr, weight_decay=config.weight_decay)        # train.py:210-228

    train_tf, _ = get_transform(config.transform_type, (H, W), {"normalization": "piano-roll", "seed": config.seed + rank,
                                                                "max_pitch_shift": config.max_pitch_shift,
                                                                "binarize_threshold": config.binarize_threshold})
    rolls = _load_rolls(config, H, W, device)
    data = DeviceRollDataset(rolls, train_tf, config.batch_size_per_gpu, rank, world, shuffle=True, seed=config.seed)
    scheduler = tv.optim.FusedOneCycleLR(optimizer, [g["lr"] for g in optimizer.param_groups], epochs=config.epochs,
                                         steps_per_epoch=len(data))                          # train.py:232-238
    # gradient sink for every world size (local at one GPU: same gradient path, no collective)
    buckets = GradientBuckets(model)
    buckets.broadcast_parameters(0)
    buckets.broadcast_buffers(0)

    start_epoch, total_step, n_samples_seen = 1, 0, 0
    if config.checkpoint_path and os.path.isfile(config.checkpoint_path):
        ckpt = torch.load(config.checkpoint_path, map_location=device, weights_only=False)
        say(f"Loading state from checkpoint (epoch {ckpt['epoch']})")
        if "model" in ckpt:
            model.load_state_dict(ckpt["model"])
        model.encoder.load_state_dict(ckpt["encoder"])                                       # train.py:325-328
        model.decoder.load_state_dict(ckpt["decoder"])
        optimizer.load_state_dict(ckpt["optimizer"])
        if ckpt.get("scheduler") is not None:
            scheduler.load_state_dict(ckpt["scheduler"])
        start_epoch, total_step, n_samples_seen = ckpt["epoch"] + 1, ckpt["total_step"], ckpt["n_samples_seen"]
    elif not config.checkpoint_path and config.models_dir:
        config.checkpoint_path = os.path.join(config.models_dir, "checkpoint_latest.pt")     # train.py:300

    step = TrainStep(model, optimizer, scheduler, buckets, use_graph=not config.no_graph)
    model.train()
    losses_host = torch.zeros(3, dtype=torch.float32).pin_memory()
    last = {}
    done = False
    for epoch in range(start_epoch, config.epochs + 1):
        t0 = time.time()
        seen0 = n_samples_seen
        for batch_idx, x in enumerate(data.epoch(epoch)):
            losses = step(x)
            total_step += 1
            n_samples_seen += config.batch_size
            if batch_idx % config.log_interval == 0 or batch_idx == len(data) - 1:
                losses_host.copy_(torch.stack([losses["loss"].reshape(()), losses["reconstruction_loss"].reshape(()),
                                               losses["kld_loss"].reshape(())]), non_blocking=False)
                last = {"loss": float(losses_host[0]), "reconstruction_loss": float(losses_host[1]), "kld_loss": float(losses_host[2])}
                say(f"Train Epoch:{epoch:4d} [{batch_idx:5d}/{len(data)}] Loss: {last['loss']:.6f}  Recon: "
                    f"{last['reconstruction_loss']:.6f}  KLD: {last['kld_loss']:.6f}  LR: {scheduler.get_last_lr()[0]:.3e}")
            if config.max_steps is not None and total_step >= config.max_steps:
                done = True
                break
        torch.cuda.synchronize()
        dt = time.time() - t0
        say(f"Epoch {epoch}/{config.epochs}: {(n_samples_seen - seen0) / max(dt, 1e-9):,.0f} samples/s over {world} GPU(s)")
        if config.checkpoint_path and rank == 0:
            save_checkpoint(config.checkpoint_path, model, optimizer, scheduler, config, epoch=epoch, total_step=total_step,
                            n_samples_seen=n_samples_seen)
            say(f"Saved model to  {config.checkpoint_path}")
        if done:
            break
    if world > 1:
        dist.barrier()
    return {"total_step": total_step, "n_samples_seen": n_samples_seen, **last}


def main(argv: Optional[list] = None) -> int:
    config = get_parser().parse_args(argv)
    run(config)
    return 0


if __name__ == "__main__":
    sys.exit(main())

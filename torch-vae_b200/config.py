"""The reference's ``configs/vae.yaml`` surface (midi_autoencoder/configs/vae.yaml:1-34).

The reference ships that PyTorch-VAE-style file but no code reads it (SURVEY.md
§2 #11); north_star asks to keep the YAML surface, so this loader accepts the
same schema and maps it onto the drop-in model and the reference's optimizer
construction (train.py:201-228):

  model_params{name,in_channels,latent_dim,hidden_dims} -> VanillaVAE ctor
  exp_params{learning_rate,weight_decay,kld_weight,manual_seed}
  data_params{train_batch_size,patch_size,...}, trainer_params, logging_params: kept verbatim
"""

from __future__ import annotations

from typing import Any, Dict, Optional

import torch
import yaml

DEFAULTS: Dict[str, Dict[str, Any]] = {
    "model_params": {"name": "VanillaVAE", "in_channels": 1, "latent_dim": 10, "hidden_dims": [32, 64, 128, 256]},
    "data_params": {"train_batch_size": 100, "val_batch_size": 100, "patch_size": 64, "num_workers": 4},
    "exp_params": {"learning_rate": 0.001, "weight_decay": 0.00001, "scheduler_gamma": 0.1, "kld_weight": 0.00025, "manual_seed": 0},
    "trainer_params": {},
    "logging_params": {},
}


def load_config(path: str) -> Dict[str, Dict[str, Any]]:
    with open(path) as f:
        raw = yaml.safe_load(f) or {}
    cfg = {k: dict(v) for k, v in DEFAULTS.items()}
    for section, values in raw.items():
        if section not in cfg:
            raise ValueError(f"unknown config section '{section}' (expected one of {sorted(cfg)})")
        cfg[section].update(values or {})
    if cfg["model_params"]["name"] != "VanillaVAE":
        raise NotImplementedError(f"model '{cfg['model_params']['name']}' is not supported")
    return cfg


def build_model(cfg: Dict[str, Dict[str, Any]], input_dim, compute_dtype: torch.dtype = torch.float32,
                device: Optional[str] = None):
    """VanillaVAE from ``model_params`` / ``exp_params`` (seeded with ``manual_seed``)."""
    from .models import VanillaVAE

    mp, ep = cfg["model_params"], cfg["exp_params"]
    torch.manual_seed(int(ep.get("manual_seed", 0)))
    model = VanillaVAE(int(mp["in_channels"]), int(mp["latent_dim"]), input_dim, hidden_dims=list(mp["hidden_dims"]),
                       kld_weight=float(ep["kld_weight"]), compute_dtype=compute_dtype)
    return model.to(device) if device is not None else model


def build_optimizer(cfg: Dict[str, Dict[str, Any]], model):
    """AdamW over the reference's two parameter groups, encoder and decoder
    (train.py:210-228), with the YAML's learning rate and weight decay."""
    ep = cfg["exp_params"]
    lr, wd = float(ep["learning_rate"]), float(ep["weight_decay"])
    groups = [
        {"params": model.encoder.parameters(), "lr": lr, "name": "encoder"},
        {"params": model.decoder.parameters(), "lr": lr, "name": "decoder"},
    ]
    return torch.optim.AdamW(groups, lr=lr, weight_decay=wd)

"""Host-side logic that needs no GPU: YAML surface, transform registry, and the
data-parallel gradient buckets over a 2-rank gloo group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_yaml_surface_matches_reference_schema(tmp_path):
    import torch_vae_b200 as tv

    text = """
model_params:
  name: 'VanillaVAE'
  in_channels: 1
  latent_dim: 10
  hidden_dims: [32, 64, 128, 256]
exp_params:
  learning_rate: 0.001
  weight_decay: 0.00001
  kld_weight: 0.00025
  manual_seed: 0
data_params:
  train_batch_size: 100
"""
    p = tmp_path / "vae.yaml"
    p.write_text(text)
    cfg = tv.config.load_config(str(p))
    model = tv.config.build_model(cfg, 32)
    assert model.kld_weight == 0.00025 and model.latent_dim == 10 and model.flattened_size == 1024
    opt = tv.config.build_optimizer(cfg, model)
    assert [g["name"] for g in opt.param_groups] == ["encoder", "decoder"]
    assert opt.param_groups[0]["lr"] == 0.001 and opt.param_groups[0]["weight_decay"] == 0.00001
    n_opt = sum(p.numel() for g in opt.param_groups for p in g["params"])
    n_enc_dec = sum(p.numel() for p in model.encoder.parameters()) + sum(p.numel() for p in model.decoder.parameters())
    assert n_opt == n_enc_dec  # fc_mu / fc_var / decoder_input / final_layer are NOT optimised (train.py:210-228)
    (tmp_path / "bad.yaml").write_text("nonsense_params: {a: 1}\n")
    with pytest.raises(ValueError):
        tv.config.load_config(str(tmp_path / "bad.yaml"))


def test_launcher_flags_follow_the_reference_cli(tmp_path):
    """torch_vae_b200.train: the train.py:770-1063 flag subset of the hot path, the configs/vae.yaml defaults, and
    the lr rule of train.py:201 (per 128 samples, scaled by the global batch)."""
    from torch_vae_b200 import train as launcher

    cfg = launcher.get_parser().parse_args([])
    launcher._apply_yaml(cfg)
    # reference defaults: train.py:877 (n_features 10), :886 (epochs 5), :894 (lr 0.01), :930 (wd 0), batch 128
    assert (cfg.n_features, cfg.epochs, cfg.lr_relative, cfg.weight_decay, cfg.batch_size_per_gpu) == (10, 5, 0.01, 0.0, 128)
    assert cfg.hidden_dims == [32, 64, 128, 256] and cfg.kld_weight == 1.0 and cfg.lr_absolute is None
    assert cfg.optimizer == "AdamW" and cfg.scheduler == "OneCycle" and cfg.arch == "VanillaVAE"
    y = tmp_path / "vae.yaml"
    y.write_text("model_params:\n  latent_dim: 24\n  hidden_dims: [64, 128, 256, 512]\nexp_params:\n  learning_rate: 0.002\n"
                 "  weight_decay: 0.0001\n  kld_weight: 0.00025\n  manual_seed: 7\ndata_params:\n  train_batch_size: 100\n"
                 "trainer_params:\n  max_epochs: 100\nlogging_params:\n  log_interval: 25\n")
    cfg = launcher.get_parser().parse_args(["--config", str(y), "--epochs", "3", "--wd", "0.5"])
    launcher._apply_yaml(cfg)
    assert (cfg.n_features, cfg.hidden_dims, cfg.kld_weight, cfg.seed) == (24, [64, 128, 256, 512], 0.00025, 7)
    assert (cfg.epochs, cfg.weight_decay, cfg.batch_size_per_gpu, cfg.log_interval) == (3, 0.5, 100, 25)   # flags win over YAML
    assert cfg.lr_absolute == 0.002   # the YAML learning rate is absolute; --lr is relative
    cfg = launcher.get_parser().parse_args(["--config", str(y), "--lr", "0.02", "--image-size", "128x1024"])
    launcher._apply_yaml(cfg)
    assert cfg.lr_absolute is None and cfg.lr_relative == 0.02 and launcher._image_size(cfg.image_size) == (128, 1024)
    assert launcher._image_size("64") == (64, 64)
    with pytest.raises(RuntimeError):   # no CUDA device here: the launcher refuses instead of falling back
        launcher.run(launcher.get_parser().parse_args([]))
    from torch_vae_b200.trainer import DeviceRollDataset

    with pytest.raises(RuntimeError):
        DeviceRollDataset(torch.zeros(8, 128, 144, dtype=torch.uint8), None, 4)   # host tensor: the dataset lives in HBM


def test_transform_registry_mirrors_reference():
    import torch_vae_b200 as tv

    dt = tv.data_transformations
    assert dt.NORMALIZATION["mnist"] == [(0.5,), (1.0,)]
    train, test = dt.get_transform("midi", 128, {"binarize_threshold": 0.5, "max_pitch_shift": 6})
    assert train.random_crop and not test.random_crop and train.max_pitch_shift == 6 and test.max_pitch_shift == 0
    assert (train.mean, train.std) == (0.5, 1.0)
    tr2, _ = dt.get_transform("noaug", (128, 1024), {"normalization": "piano-roll"})
    assert (tr2.height, tr2.width, tr2.mean) == (128, 1024, 0.0)
    with pytest.raises(NotImplementedError):
        dt.get_transform("unknown")
    with pytest.raises(RuntimeError):
        train(torch.zeros(2, 128, 160, dtype=torch.uint8))  # CPU tensor: no fallback


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dp_worker(rank, world, port, q):
    import sys

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch_vae_b200 as tv
    from torch_vae_b200.parallel import GradientBuckets, execution_order_reversed

    torch.manual_seed(0)
    model = tv.VanillaVAE(1, 6, 32, hidden_dims=[8, 16, 32, 64])
    buckets = GradientBuckets(model, bucket_cap_mb=0.02, first_bucket_cap_mb=0.001)
    assert len(buckets.buckets) >= 3 and model._grad_sink is buckets
    order = execution_order_reversed(model)
    assert len(order) == 40 and order[0] is model.final_layer[3].weight and order[-1] is model.encoder[0][1].bias
    # every rank "produces" rank-dependent gradients in backward order, as engine.py does
    for step in range(2):
        buckets.begin()
        for k, p in enumerate(order):
            g = torch.full_like(p, float(rank + 1) * (k + 1) * (step + 1))
            buckets.push(p, g)
        buckets.finish()
        mean_scale = sum(r + 1 for r in range(world)) / world
        for k, p in enumerate(order):
            expect = mean_scale * (k + 1) * (step + 1)
            assert torch.allclose(p.grad, torch.full_like(p, expect)), (k, float(p.grad.flatten()[0]), expect)
    # an incomplete step must be reported, not silently reduced
    buckets.begin()
    buckets.push(order[0], torch.zeros_like(order[0]))
    try:
        buckets.finish()
        ok = False
    except RuntimeError:
        ok = True
    # buffers follow rank 0
    model.encoder[0][1].running_mean.fill_(float(rank))
    buckets.broadcast_buffers(0)
    ok = ok and float(model.encoder[0][1].running_mean[0]) == 0.0
    q.put((rank, ok))
    dist.destroy_process_group()


def test_inference_mode_rule():
    """engine.inference_mode: the conv epilogues may apply eval-mode BatchNorm + LeakyReLU (one pass, no stored
    pre-normalisation tensor) only when the module is in eval mode AND autograd will record nothing."""
    import torch_vae_b200 as tv

    im = tv.engine.inference_mode
    w = torch.zeros(3, requires_grad=True)
    x = torch.zeros(3)
    assert not im(True, x, w)                      # training: batch statistics, activations are kept
    assert not im(False, x, w)                     # eval, but a parameter requires grad and grad mode is on
    assert im(False, x, w.detach())                # eval, nothing requires grad
    assert im(False, x, None, w.detach())          # None entries (absent biases) are ignored
    with torch.no_grad():
        assert im(False, x, w)                     # eval under no_grad: sample(), evaluate()
        assert not im(True, x, w)
    assert not im(False, x.requires_grad_(True), w.detach())   # gradient w.r.t. the input (e.g. latent optimisation)


def test_gradient_buckets_local_sink_without_process_group():
    """Without torch.distributed the sink is local (one GPU): same slots and views, no collective, p.grad = the bucket
    views holding exactly what was pushed; an incomplete step is still reported."""
    import torch_vae_b200 as tv
    from torch_vae_b200.parallel import GradientBuckets, execution_order_reversed

    assert not dist.is_initialized()
    torch.manual_seed(0)
    model = tv.VanillaVAE(1, 6, 32, hidden_dims=[8, 16, 32, 64])
    buckets = GradientBuckets(model, bucket_cap_mb=0.02, first_bucket_cap_mb=0.001)
    assert buckets.local and buckets.world == 1 and buckets.comm_stream is None and model._grad_sink is buckets
    order = execution_order_reversed(model)
    for step in range(2):
        buckets.begin()
        for k, p in enumerate(order):
            if k % 2:   # written in place by a "kernel" (engine._dst) ...
                buckets.view(p).fill_(float((k + 1) * (step + 1)))
                buckets.push(p, buckets.view(p))
            else:       # ... or produced elsewhere and staged
                buckets.push(p, torch.full_like(p, float((k + 1) * (step + 1))))
        buckets.finish()
        for k, p in enumerate(order):
            assert torch.equal(p.grad, torch.full_like(p, float((k + 1) * (step + 1))))   # overwritten, not accumulated
            assert p.grad.data_ptr() == buckets.view(p).data_ptr()
    buckets.broadcast_parameters(0)   # no-ops locally
    buckets.broadcast_buffers(0)
    buckets.begin()
    buckets.push(order[0], torch.zeros_like(order[0]))
    with pytest.raises(RuntimeError):
        buckets.finish()


def test_gradient_buckets_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    results = sorted(q.get(timeout=5) for _ in range(2))
    assert results == [(0, True), (1, True)]

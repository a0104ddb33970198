"""The drop-in against the reference's OWN host code:

  * bit-identical initial weights from the same ``torch.manual_seed`` (a2, models.py:227-247);
  * ``evaluate()`` (evaluation.py:12-113) run unmodified against the drop-in;
  * ``safe_save_model`` (utils.py:311-351) -> ``torch.load`` -> ``load_state_dict`` round trip with the
    checkpoint contents of train.py:445-460 (encoder, decoder, optimizer, scheduler);
  * FusedAdamW + FusedOneCycleLR under CUDA-graph replay == torch.optim.AdamW + OneCycleLR eager
    (train.py:210-238,656-659);
  * a checkpoint written by the stock torch.optim.AdamW loads into FusedAdamW.

The reference modules come from ``baseline/_ref`` (installed from /root/reference in the build
container, shipped to the GPU box); tests that need them skip when it is absent.
"""
import os

import numpy as np
import pytest
import torch

from oracle import reference_driver as rd
from oracle import vae_oracle as vo

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not rd.available(), reason="baseline/_ref (the installed reference) is not present")


@pytest.fixture(scope="module")
def tv():
    import torch_vae_b200 as m

    m.load_library()
    return m


@needs_ref
def test_same_seed_gives_bit_identical_initial_weights(tv):
    """a2: same constructor order and init rules as models.py:40-83,227-247 (ConvTranspose2d keeps torch's
    default init, the three Linear layers are never passed to _init_weights)."""
    torch.manual_seed(0)
    ref = rd.build_model(32, 32, [32, 64, 128, 256], 10)
    torch.manual_seed(0)
    got = tv.VanillaVAE(1, 10, 32)
    sd_ref, sd_got = ref.state_dict(), got.state_dict()
    assert list(sd_ref) == list(sd_got) and len(sd_ref) == 64
    for k in sd_ref:
        assert torch.equal(sd_ref[k], sd_got[k]), k
    assert got.hidden_dims == ref.hidden_dims == [256, 128, 64, 32]  # reversed in place (models.py:60)
    assert str(got.encoder) == str(ref.encoder) and str(got.final_layer) == str(ref.final_layer)
    # generalised extents: the encoder is drawn before any Linear layer, so it sees the same random stream as the
    # reference class; the three Linear layers have another size there (models.py:33 literal), consume another
    # number of draws, and everything constructed after them differs from the 32x32-literal class by construction
    torch.manual_seed(7)
    ref = rd.build_model(128, 128, [32, 64, 128, 256], 10)
    torch.manual_seed(7)
    got = tv.VanillaVAE(1, 10, (128, 128))
    for k, v in ref.state_dict().items():
        if k.startswith("encoder."):
            assert torch.equal(v, got.state_dict()[k]), k
        else:
            assert v.shape == got.state_dict()[k].shape or k.startswith(("fc_", "decoder_input")), k


class _Rolls(torch.utils.data.Dataset):
    def __init__(self, n, H, W):
        g = torch.Generator().manual_seed(4)
        self.x = (torch.rand(n, 1, H, W, generator=g) < 0.1).float()

    def __len__(self):
        return len(self.x)

    def __getitem__(self, i):
        return self.x[i], 0


@needs_ref
def test_reference_evaluate_runs_against_the_drop_in(tv):
    """evaluation.py:42-76 calls model.eval(), model(stimuli) under no_grad and reads output / latents."""
    evaluate = rd.import_module("evaluation").evaluate
    H = W = 32
    params = vo.closed_form_params(vo.param_shapes(1, 10, H, W, [32, 64, 128, 256]))
    model = tv.VanillaVAE(1, 10, 32)
    model.load_state_dict({**model.state_dict(), **params})
    model = model.cuda()
    ref = rd.build_model(H, W, [32, 64, 128, 256], 10)
    ref.load_state_dict({**ref.state_dict(), **params})
    loader = torch.utils.data.DataLoader(_Rolls(24, H, W), batch_size=8)
    # reparameterize samples in eval mode too (models.py:181-183): compare the mean reconstruction error, which
    # depends on the latents only through a [B,10] noise vector, over the same data
    torch.manual_seed(0)
    got = evaluate(loader, model, torch.device("cuda"), verbosity=0)
    torch.manual_seed(0)
    want = evaluate(loader, ref, torch.device("cpu"), verbosity=0)
    assert got["count"] == want["count"] == 24
    assert got["cross-entropy"] == want["cross-entropy"] == 0.0  # C == 1: always zero
    assert abs(got["mse"] - want["mse"]) <= 0.05 * want["mse"]
    assert abs(got["mae"] - want["mae"]) <= 0.05 * want["mae"]
    assert not model.training


@needs_ref
def test_reference_checkpoint_round_trip(tv, tmp_path):
    """utils.safe_save_model with the contents of train.py:445-460, reloaded the way train.py:75,320-329 does."""
    safe_save_model = rd.import_module("utils").safe_save_model
    torch.manual_seed(1)
    model = tv.VanillaVAE(1, 10, 32).cuda().train()
    groups = [{"params": model.encoder.parameters(), "lr": 1e-3, "name": "encoder"},
              {"params": model.decoder.parameters(), "lr": 1e-3, "name": "decoder"}]
    opt = tv.FusedAdamW(groups, lr=1e-3, weight_decay=0.0)
    sched = tv.optim.FusedOneCycleLR(opt, [1e-3, 1e-3], epochs=2, steps_per_epoch=10)
    x = _Rolls(8, 32, 32).x.cuda()
    for _ in range(3):
        out = model(x)
        opt.zero_grad()
        model.loss(out)["loss"].backward()
        opt.step()
        sched.step()
    path = os.path.join(tmp_path, "ckpt", "checkpoint.pt")
    safe_save_model({"encoder": model.encoder, "decoder": model.decoder, "optimizer": opt, "scheduler": sched},
                    path, config=None, epoch=1, total_step=3, n_samples_seen=24)
    assert os.path.isfile(path) and not os.path.exists(os.path.join(tmp_path, "ckpt", ".tmp.checkpoint.pt"))
    ck = torch.load(path, map_location="cuda", weights_only=False)
    assert ck["epoch"] == 1 and ck["total_step"] == 3
    # (a) into a fresh drop-in
    torch.manual_seed(2)
    m2 = tv.VanillaVAE(1, 10, 32).cuda().train()
    m2.encoder.load_state_dict(ck["encoder"])
    m2.decoder.load_state_dict(ck["decoder"])
    o2 = tv.FusedAdamW([{"params": m2.encoder.parameters(), "lr": 1e-3}, {"params": m2.decoder.parameters(), "lr": 1e-3}],
                       lr=1e-3, weight_decay=0.0)
    o2.load_state_dict(ck["optimizer"])
    s2 = tv.optim.FusedOneCycleLR(o2, [1e-3, 1e-3], epochs=2, steps_per_epoch=10)
    s2.load_state_dict(ck["scheduler"])
    for a, b in zip(model.encoder.state_dict().values(), m2.encoder.state_dict().values()):
        assert torch.equal(a, b)
    assert s2.device_step_count() == 3 and np.allclose(s2.get_last_lr(), sched.get_last_lr(), rtol=1e-6)
    p_a, p_b = next(model.decoder.parameters()), next(m2.decoder.parameters())
    assert torch.equal(opt.state[p_a]["exp_avg"], o2.state[p_b]["exp_avg"])
    # (b) into the reference class itself: same keys, same shapes
    ref = rd.build_model(32, 32, [32, 64, 128, 256], 10)
    ref.encoder.load_state_dict(ck["encoder"])
    ref.decoder.load_state_dict(ck["decoder"])
    ref_opt = torch.optim.AdamW([{"params": ref.encoder.parameters()}, {"params": ref.decoder.parameters()}], lr=1e-3)
    ref_opt.load_state_dict(torch.load(path, map_location="cpu", weights_only=False)["optimizer"])


def test_stock_adamw_checkpoint_loads_into_fused(tv):
    """A checkpoint written by the reference's own optimizer (capturable=False, CPU step) must load."""
    torch.manual_seed(3)
    ps = [torch.randn(16, 8, requires_grad=True), torch.randn(8, requires_grad=True)]
    ref = torch.optim.AdamW([{"params": ps}], lr=1e-3)
    for p in ps:
        p.grad = torch.randn_like(p)
    ref.step()
    ref.step()
    sd = ref.state_dict()  # CPU tensors, step on the CPU
    qs = [p.detach().clone().cuda().requires_grad_(True) for p in ps]
    got = tv.FusedAdamW([{"params": qs}], lr=1e-3)
    got.load_state_dict(sd)
    for q in qs:
        assert got.state[q]["step"].is_cuda and float(got.state[q]["step"]) == 2.0
        q.grad = torch.ones_like(q)
    got.step()
    torch.cuda.synchronize()
    assert float(got.state[qs[0]]["step"]) == 3.0
    assert got.param_groups[0]["capturable"] is True


def test_fused_adamw_follows_host_scheduler_eagerly(tv):
    """torch's own OneCycleLR rewrites param_groups on the host; FusedAdamW re-uploads its table."""
    torch.manual_seed(5)
    shapes = [(32, 1, 3, 3), (32,), (64, 32, 3, 3)]
    ref_p = [torch.randn(s, device="cuda").requires_grad_(True) for s in shapes]
    got_p = [p.detach().clone().requires_grad_(True) for p in ref_p]
    mk = lambda ps: [{"params": ps[:2], "lr": 3e-3}, {"params": ps[2:], "lr": 1e-3}]
    ref = torch.optim.AdamW(mk(ref_p), lr=1e-3, weight_decay=0.01)
    got = tv.FusedAdamW(mk(got_p), lr=1e-3, weight_decay=0.01)
    s_ref = torch.optim.lr_scheduler.OneCycleLR(ref, [3e-3, 1e-3], total_steps=12)
    s_got = torch.optim.lr_scheduler.OneCycleLR(got, [3e-3, 1e-3], total_steps=12)
    for it in range(10):
        for a, b in zip(ref_p, got_p):
            g = torch.randn_like(a)
            a.grad, b.grad = g.clone(), g.clone()
        ref.step(); s_ref.step()
        got.step(); s_got.step()
    for a, b in zip(ref_p, got_p):
        assert vo.rel_err(b, a) <= 3e-6


def test_fused_onecycle_under_graph_replay_matches_torch(tv):
    """20 CUDA-graph replays of (FusedAdamW.step, FusedOneCycleLR.step) == torch.optim.AdamW + OneCycleLR eager:
    lr and beta1 change every step although the launch arguments were frozen at capture (train.py:232-238,659)."""
    torch.manual_seed(6)
    shapes = [(32, 1, 3, 3), (32,), (64, 32, 3, 3), (64,)]
    ref_p = [torch.randn(s, device="cuda").requires_grad_(True) for s in shapes]
    got_p = [p.detach().clone().requires_grad_(True) for p in ref_p]
    mk = lambda ps: [{"params": ps[:2], "lr": 3e-3, "name": "encoder"}, {"params": ps[2:], "lr": 1e-3, "name": "decoder"}]
    ref = torch.optim.AdamW(mk(ref_p), lr=1e-3, weight_decay=0.01)
    got = tv.FusedAdamW(mk(got_p), lr=1e-3, weight_decay=0.01)
    total = 40
    s_ref = torch.optim.lr_scheduler.OneCycleLR(ref, [g["lr"] for g in ref.param_groups], epochs=4, steps_per_epoch=10)
    s_got = tv.optim.FusedOneCycleLR(got, [g["lr"] for g in got.param_groups], epochs=4, steps_per_epoch=10)
    grads = [torch.randn(total, *s, device="cuda") for s in shapes]
    static_g = [torch.zeros(s, device="cuda") for s in shapes]
    for b, g in zip(got_p, static_g):
        b.grad = g
    lrs_ref, lrs_got = [], []

    def ref_step(i):
        for a, g in zip(ref_p, grads):
            a.grad = g[i].clone()
        ref.step(); s_ref.step()
        lrs_ref.append([grp["lr"] for grp in ref.param_groups] + [ref.param_groups[0]["betas"][0]])

    # three eager steps, then capture, then replays
    for i in range(3):
        ref_step(i)
        for s_, g in zip(static_g, grads):
            s_.copy_(g[i])
        got.step(); s_got.step()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph):
            got.step(); s_got.step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    # the capture itself does not execute: replay for steps 3 .. 22
    for i in range(3, 23):
        ref_step(i)
        for s_, g in zip(static_g, grads):
            s_.copy_(g[i])
        graph.replay()
        lrs_got.append(s_got.get_last_lr() + [float(got.hyper_table("cuda")[0, 1])])
    torch.cuda.synchronize()
    assert s_got.device_step_count() == 23
    for r, g in zip(lrs_ref[3:], lrs_got):
        assert np.allclose(r, g, rtol=2e-6), (r, g)
    assert abs(lrs_ref[-1][0] - lrs_ref[3][0]) > 1e-4  # the schedule really moved
    for a, b in zip(ref_p, got_p):
        assert vo.rel_err(b, a) <= 5e-6


def test_launcher_trains_checkpoints_and_resumes(tv, tmp_path):
    """python -m torch_vae_b200.train: device-resident rolls, graph-replayed step, the safe_save_model file layout
    (utils.py:311-351 / train.py:445-460) and a resume that continues the counters and the OneCycle schedule."""
    from torch_vae_b200 import train as launcher

    ck = str(tmp_path / "models" / "checkpoint_latest.pt")
    argv = ["--image-size", "32", "--batch-size", "16", "--n-rolls", "128", "--epochs", "2", "--checkpoint", ck, "--seed", "3",
            "--log-interval", "4", "--dtype", "fp32"]
    r1 = launcher.main(argv[:])
    assert r1 == 0 and os.path.isfile(ck)
    data = torch.load(ck, map_location="cpu", weights_only=False)
    assert {"encoder", "decoder", "optimizer", "scheduler", "config", "epoch", "total_step", "n_samples_seen"} <= set(data)
    assert data["epoch"] == 2 and data["total_step"] == 16 and data["n_samples_seen"] == 16 * 16
    if rd.available():   # the reference's resume path (train.py:325-328) finds its keys and shapes
        ref = rd.build_model(32, 32, [32, 64, 128, 256], 10)
        ref.encoder.load_state_dict(data["encoder"])
        ref.decoder.load_state_dict(data["decoder"])
        torch.optim.AdamW([{"params": ref.encoder.parameters()}, {"params": ref.decoder.parameters()}]).load_state_dict(data["optimizer"])
    # resume for one more epoch: counters continue
    cfg = launcher.get_parser().parse_args(argv[:7] + ["3"] + argv[8:])
    out = launcher.run(cfg)
    assert out["total_step"] == 24 and out["n_samples_seen"] == 24 * 16
    assert out["loss"] == out["loss"] and 0.0 < out["loss"] < 10.0   # finite

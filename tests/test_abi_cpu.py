"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports
every symbol include/vae_b200.h declares; the host module mirrors the reference's
surface without needing a GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vae_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vae_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import torch_vae_b200 as tv

    assert os.path.exists(tv.LIB_PATH), "build the extension first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(tv.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 23
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vae_b200.h but not exported"
    from torch_vae_b200 import _lib

    assert set(_lib.PROTOTYPES) == set(names), "ctypes prototypes out of sync with the header"
    assert tv.load_library().vae_abi_version() == 1


def test_module_surface_matches_reference_contract():
    import torch_vae_b200 as tv

    hd = [32, 64, 128, 256]
    m = tv.VanillaVAE(1, 10, 32, hidden_dims=hd)
    assert hd == [256, 128, 64, 32], "reference reverses hidden_dims in place (models.py:60)"
    assert m.name == "VanillaVAE" and m.latent_dim == 10 and m.in_channels == 1 and m.kld_weight == 1.0
    assert m.flattened_size == 1024 and m.last_conv_size == 4 and m.hidden_dims == [256, 128, 64, 32]
    sd = m.state_dict()
    assert len(sd) == 64 and len(list(m.parameters())) == 40
    assert sum(p.numel() for p in m.parameters()) == 817909
    for k, shape in {
        "encoder.0.0.weight": (32, 1, 3, 3), "encoder.3.1.running_var": (256,), "fc_mu.weight": (10, 1024),
        "fc_var.bias": (10,), "decoder_input.weight": (1024, 10), "decoder.0.0.weight": (256, 128, 3, 3),
        "decoder.2.1.num_batches_tracked": (), "final_layer.0.weight": (32, 32, 3, 3), "final_layer.3.weight": (1, 32, 3, 3),
    }.items():
        assert tuple(sd[k].shape) == shape, k
    assert sum(p.numel() for p in tv.VanillaVAE(1, 10, 128).parameters()) == 1294069
    for name in ("encode", "decode", "reparameterize", "forward", "loss", "sample", "generate"):
        assert callable(getattr(m, name))
    # the optimizer groups the reference builds (train.py:210-228) must be constructible
    torch.optim.AdamW([{"params": m.encoder.parameters()}, {"params": m.decoder.parameters()}], lr=1e-3)
    m.encoder.load_state_dict(m.encoder.state_dict())  # train.py:325


def test_no_cpu_fallback():
    import torch_vae_b200 as tv

    m = tv.VanillaVAE(1, 10, 32)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(2, 1, 32, 32))
    with pytest.raises(RuntimeError):
        tv.ops.sigmoid(torch.zeros(4))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "torch-vae_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("oracle/", ""), f"{fn} mentions the oracle"

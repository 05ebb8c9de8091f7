"""Shape / parameter-count / wrapper checks of the PyTorch denoiser used for the separate UNet timing (CPU, tiny config)."""
import torch

from free_hunch_b200.denoiser import ADMUNet, CONFIGS, IDDPMLinearPrecond


def test_tiny_unet_shapes_and_vjp():
    net = ADMUNet(image_size=32, model_channels=32, num_res_blocks=1, attention_resolutions=(8,), channel_mult=(1, 2, 2), head_channels=16)
    den = IDDPMLinearPrecond(net, use_fp16=False)
    x = torch.randn(2, 3, 32, 32, dtype=torch.float64, requires_grad=True)
    d, var = den(x, torch.tensor(5.0, dtype=torch.float64))
    assert d.shape == x.shape and d.dtype == torch.float64 and var is None          # float64 through c_out
    assert float(d.abs().max()) <= 1.0
    (g,) = torch.autograd.grad((d * torch.randn_like(d)).sum(), x)
    assert g.shape == x.shape and torch.isfinite(g).all()


def test_timestep_index_matches_sigma_grid():
    den = IDDPMLinearPrecond(torch.nn.Identity())
    idx = den.round_sigma_index(torch.tensor([80.2878, 0.010001], dtype=torch.float64))
    assert abs(float(den.u[idx[0]]) - 80.2878) < 1e-3 and abs(float(den.u[idx[1]]) - 0.010001) < 1e-5


def test_imagenet256_parameter_count():
    with torch.device("meta"):
        net = ADMUNet(**CONFIGS["imagenet256"])
    n = sum(p.numel() for p in net.parameters())
    assert 5.3e8 < n < 5.7e8, n                    # guided-diffusion's 256x256 unconditional model has ~552 M parameters


def _unet_golden():
    import os
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "unet_small.npz"))
    sd = {k[3:]: (torch.from_numpy(g[k].astype("int32")) << 16).view(torch.float32) for k in g.files if k.startswith("sd/")}
    # the generator's CFG (tests/golden/make_golden_unet.py): attention at down-sampling rates 2 and 4 of a 32 x 32 input
    net = ADMUNet(image_size=32, model_channels=32, num_res_blocks=2, attention_resolutions=(16, 8), channel_mult=(1, 1, 2),
                  head_channels=16).eval()
    return g, net.load_reference_state_dict(sd)


def test_unet_matches_reference_unetmodel():
    """ADMUNet fed the state dict of the REAL reference's UNetModel (training/openai_unet.py:395-686, golden produced by
    tests/golden/make_golden_unet.py) reproduces its output: same architecture, same parameter layout (incl. the
    legacy-order attention), float32 rounding apart."""
    g, net = _unet_golden()
    assert sum(p.numel() for p in net.parameters()) == sum(g[k].size for k in g.files if k.startswith("sd/"))
    with torch.no_grad():
        y = net(torch.from_numpy(g["x"]), torch.from_numpy(g["t"]))
    ref = torch.from_numpy(g["y"])
    assert y.shape == ref.shape == (2, 6, 32, 32)
    err = float((y - ref).norm() / ref.norm())
    assert err < 1e-5, err


def test_precond_matches_reference_wrapper():
    """IDDPMLinearPrecond against the reference's iDDPMLinearPrecond (training/openai_preconditioning.py:93-207): the
    sigma grid u, the rounded timestep index, the output dtype (float64 through c_out) and D_x at three noise levels."""
    g, net = _unet_golden()
    den = IDDPMLinearPrecond(net, use_fp16=False).eval()
    assert torch.equal(den.u, torch.from_numpy(g["u"]))
    for i in range(3):
        x, s = torch.from_numpy(g[f"pre{i}_x"]), float(g[f"pre{i}_sigma"])
        assert int(den.round_sigma_index(torch.tensor(s))[0]) == int(g[f"pre{i}_index"])
        with torch.no_grad():
            d, _ = den(x, torch.tensor(s))
        ref = torch.from_numpy(g[f"pre{i}_d"])
        assert str(d.dtype) == str(g[f"pre{i}_dtype"]) == "torch.float64"
        err = float((d - ref).norm() / ref.norm())
        assert err < 1e-5, (i, err)


def test_reference_state_dict_must_be_consumed_entirely():
    import pytest
    g, net = _unet_golden()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    with pytest.raises((KeyError, RuntimeError)):
        net.load_reference_state_dict({"time_embed.0.weight": sd["time_embed.0.weight"]})

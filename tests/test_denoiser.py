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

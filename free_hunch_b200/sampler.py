"""Sampler-side pieces around the mechanism (reference generate_conditional.py:110-201 and
training/openai_preconditioning.py:116-129,203-207): the sigma grid, sigma rounding, the Euler/Heun loop, and the
analytic stand-in denoiser used by the benchmarks when no UNet weights are available."""
import numpy as np
import torch


def iddpm_linear_u(M=1000, beta_min=1e-4, beta_max=0.02):
    """openai_preconditioning.py:116-122, evaluated with the same float32 torch ops as the reference so the grid
    is bit-identical (the cumprod near alpha_bar ~ 1 is sensitive to evaluation order)."""
    betas = torch.cat([torch.tensor([0]), torch.linspace(beta_min, beta_max, M)])
    alpha_bar = torch.cumprod(1 - betas, dim=0).flip(dims=[0])
    return torch.sqrt((1 - alpha_bar) / alpha_bar)


def round_sigma(sigma, u):
    """openai_preconditioning.py:203-207: nearest u_j (float32 cdist argmin), returned in sigma's dtype."""
    sigma = torch.as_tensor(sigma)
    index = torch.cdist(sigma.to(torch.float32).reshape(1, -1, 1), u.reshape(1, -1, 1)).argmin(2)
    return u[index.flatten()].to(sigma.dtype).reshape(sigma.shape)


def edm_sigma_steps(num_steps, sigma_min=0.002, sigma_max=80.0, rho=7.0, u=None):
    """generate_conditional.py:74-75,82,200,111-112: EDM grid clipped to the net range, rounded, 0 appended (float64)."""
    if u is None:
        u = iddpm_linear_u()
    net_min, net_max = float(u[len(u) - 2]), float(u[0])
    sigma_min, sigma_max = max(sigma_min, net_min), min(sigma_max, net_max)
    idx = torch.arange(num_steps, dtype=torch.float64)
    steps = (sigma_max ** (1 / rho) + idx / (num_steps - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
    t = round_sigma(steps, u)
    return torch.cat([t, torch.zeros_like(t[:1])])


@torch.no_grad()
def conditional_sampler(mechanism, net, y, noise, t_steps, solver="heun"):
    """generate_conditional.py:114-167 for schedule='linear', scaling='none', S_churn=0, alpha=1.
    `noise` (B,C,H,W); x is carried in float64 like the reference (:116)."""
    t_steps = [float(t) for t in t_steps]
    x_next = noise.to(torch.float64) * t_steps[0]
    n = len(t_steps) - 1
    # sigma goes to the mechanism as a 0-dim float64 HOST tensor: the mechanism and the denoiser wrapper read it with
    # .item(), which on a device tensor is a full synchronisation per call
    sig = lambda t: torch.tensor(t, dtype=torch.float64)
    fused = x_next.is_cuda                         # the step algebra as one launch each (fh_mech.cuh; bit-identical to the expressions below)
    if fused:
        from . import engine
    for i, (t_cur, t_next) in enumerate(zip(t_steps[:-1], t_steps[1:])):
        x_hat = x_next
        h = t_next - t_cur
        with torch.enable_grad():
            den = mechanism(x_hat.detach().clone().requires_grad_(), net, y, sig(t_cur)).detach()
        if fused:
            d_cur, x_prime = engine.sampler_euler(x_hat.contiguous(), den.contiguous(), t_cur, h)
        else:
            d_cur = (x_hat - den) / t_cur
            x_prime = x_hat + h * d_cur
        if solver == "euler" or i == n - 1:
            x_next = x_prime if fused else x_hat + h * d_cur
        else:
            with torch.enable_grad():
                den2 = mechanism(x_prime.detach().clone().requires_grad_(), net, y, sig(t_next)).detach()
            if fused:
                x_next = engine.sampler_heun(x_hat.contiguous(), d_cur, x_prime, den2.contiguous(), t_next, h)
            else:
                d_prime = x_prime / t_next - den2 / t_next
                x_next = x_hat + h * (0.5 * d_cur + 0.5 * d_prime)
    if hasattr(mechanism, "finalize"):
        mechanism.finalize()                       # deferred host reads (one synchronisation), no-op otherwise
    return x_next


class _SymmetricDctGain(torch.autograd.Function):
    """y = IDCT2(gain .* DCT2(x)): linear and symmetric, so the VJP is the same map."""

    @staticmethod
    def forward(ctx, x, gain, fh_ctx):
        ctx.gain, ctx.fh_ctx = gain, fh_ctx
        return fh_ctx.dct2(fh_ctx.dct2(x) * gain, inverse=True).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return ctx.fh_ctx.dct2(ctx.fh_ctx.dct2(g) * ctx.gain, inverse=True).to(g.dtype), None, None


class AnalyticDCTGaussianDenoiser(torch.nn.Module):
    """D(x, sigma) = IDCT2( v/(v+sigma^2) .* DCT2(x) ): the exact denoiser of a DCT-diagonal Gaussian prior with
    variances v.  Stand-in for the UNet (its weights are not distributed with the
    reference); differentiable, returns (D_x, x0_var) like the reference's preconditioning wrapper."""

    def __init__(self, dct_variance, fh_ctx):
        super().__init__()
        self.fh_ctx = fh_ctx
        self.register_buffer("v", torch.as_tensor(dct_variance, dtype=torch.float32).to(fh_ctx.device))

    def forward(self, x, sigma):
        s2 = float(sigma) ** 2
        gain = self.v / (self.v + s2)
        return _SymmetricDctGain.apply(x, gain, self.fh_ctx), None


class AnalyticDCTGaussianDenoiser64(torch.nn.Module):
    """The same stand-in denoiser in float64 PyTorch (dense 256 x 256 DCT matrices, plain matmuls, autograd VJP): what
    the golden runs of the reference use as `model` (tests/golden/make_golden.py), so that a parity run differs from the
    reference only in the hot path under test and not in the stand-in's own float32 transforms."""

    def __init__(self, dct_variance, device):
        super().__init__()
        n = dct_variance.shape[-1]
        kk = np.arange(n)[:, None]
        nn = np.arange(n)[None, :]
        cm = np.sqrt(2.0 / n) * np.cos(np.pi * (2 * nn + 1) * kk / (2 * n))
        cm[0] /= np.sqrt(2.0)
        self.register_buffer("cm", torch.from_numpy(cm).to(device))
        self.register_buffer("v", torch.as_tensor(np.asarray(dct_variance), dtype=torch.float64).to(device))

    def forward(self, x, sigma):
        gain = self.v / (self.v + float(sigma) ** 2)
        xh = self.cm @ x.to(torch.float64) @ self.cm.T
        return self.cm.T @ (gain * xh) @ self.cm, None

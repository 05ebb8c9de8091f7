"""Measurement operators A / A^T: host-side mirror of measurement_utils/measurements.py (reference :37-320)
for the four linear operators Free Hunch is benchmarked on.  forward / transpose run the direct pixel-space
CUDA kernels (circular convolution with the PSF, x4 decimation / zero insertion, mask multiply); the OTF
tuple `pre_calculated = (FB, FBC, F2B, FBFy)` the reference caches is provided for API compatibility only and
is not used by the product solvers.
"""
import os

import numpy as np
import torch

from . import engine
from .covariance import get_context, _default_device

__OPERATOR__ = {}


def register_operator(name):
    def wrapper(cls):
        if __OPERATOR__.get(name, None):
            raise NameError(f"Name {name} is already registered!")
        cls.name = name
        __OPERATOR__[name] = cls
        return cls
    return wrapper


def get_operator(name, **kwargs):
    """measurements.py:37-40."""
    if __OPERATOR__.get(name, None) is None:
        raise NameError(f"Name {name} is not defined.")
    return __OPERATOR__[name](**kwargs)


def p2o(psf, shape):
    """PSF -> OTF, utils_sisr.py:22-41 (zero pad, roll by -int(ks/2), fft2)."""
    otf = torch.zeros(psf.shape[:-2] + tuple(shape)).type_as(psf)
    otf[..., :psf.shape[-2], :psf.shape[-1]].copy_(psf)
    for axis, axis_size in enumerate(psf.shape[-2:]):
        otf = torch.roll(otf, -int(axis_size / 2), dims=axis + 2)
    return torch.fft.fftn(otf, dim=(-2, -1))


def _kernel_file(name, kernel_dir):
    for base in ([kernel_dir] if kernel_dir else []) + [os.path.join("measurement_utils", "kernels")]:
        p = os.path.join(base, name)
        if os.path.exists(p):
            return p
    raise FileNotFoundError(f"{name}: pass kernel=<array> or kernel_dir=<dir holding the reference's kernel files>")


class _LinearOperator:
    scale_factor = 1

    def _setup(self, in_shape, sigma_s, device):
        self.device = torch.device(device) if device is not None else _default_device()
        self.in_shape = tuple(in_shape)
        C, H, W = self.in_shape[-3:]
        self.ctx = get_context(self.device, C, H, W, max_batch=max(1, self.in_shape[0] if len(self.in_shape) == 4 else 1))
        self.sigma_s = torch.Tensor([sigma_s]).to(self.device)           # 1-element float32 tensor, like the reference
        self.pre_calculated = None

    def _grow(self, batch):
        if batch > self.ctx.max_batch:
            raise ValueError(f"operator workspace was created for batch {self.ctx.max_batch}, got {batch}")

    def A(self, x):
        self._grow(x.shape[0])
        return self.op.apply(x, transpose=False)

    def AT(self, y):
        self._grow(y.shape[0])
        return self.op.apply(y, transpose=True)


class _ConvOperator(_LinearOperator):
    def _make(self, kernel, sf=1):
        self.kernel = torch.as_tensor(np.asarray(kernel), dtype=torch.float32)     # torch.Tensor(kernel): float32
        self.op = engine.Operator.conv(self.ctx, self.kernel.cpu().numpy(), sf)

    def get_kernel(self):
        return self.kernel.view(1, 1, *self.kernel.shape).to(self.device)

    def _otf(self, y):
        C, H, W = self.in_shape[-3:]
        FB = p2o(self.get_kernel(), (H, W))
        FBC = torch.conj(FB)
        return FB, FBC, torch.pow(torch.abs(FB), 2)

    def forward(self, data, flatten=False, noiseless=False):
        dtype = data.dtype
        y = self.A(data).to(dtype)
        if not noiseless:
            y = y + self.sigma_s.to(dtype) * torch.randn_like(y)
        FB, FBC, F2B = self._otf(y)
        self.pre_calculated = (FB, FBC, F2B, None)
        if flatten:
            return y, y.reshape(y.shape[0], -1)
        return y

    def transpose(self, y, flatten=False):
        if flatten:
            y = y.reshape(y.shape[0], *self.out_shape[-3:])
        return self.AT(y).to(y.dtype)


@register_operator(name="gaussian_blur")
class GaussialBlurOperator(_ConvOperator):
    """measurements.py:164-201 (spelling of the class name kept)."""

    def __init__(self, in_shape, kernel_size=61, intensity=3.0, sigma_s=0.1, device=None, kernel=None, kernel_dir=None, **kwargs):
        self._setup(in_shape, sigma_s, device)
        self.kernel_size = kernel_size
        if kernel is None:
            kernel = np.load(_kernel_file("gaussian_ks61_std3.0.npy", kernel_dir))
        self._make(kernel)
        self.out_shape = self.in_shape


@register_operator(name="motion_blur")
class MotionBlurOperator(_ConvOperator):
    """measurements.py:126-161."""

    def __init__(self, in_shape, kernel_size=61, intensity=0.5, sigma_s=0.1, device=None, kernel=None, kernel_dir=None, **kwargs):
        self._setup(in_shape, sigma_s, device)
        self.kernel_size = kernel_size
        if kernel is None:
            kernel = np.load(_kernel_file("motion_ks61_std0.5.npy", kernel_dir))
        self._make(kernel)
        self.out_shape = self.in_shape


@register_operator(name="super_resolution")
class SuperResolutionOperator(_ConvOperator):
    """measurements.py:87-123.  A = downsample_sf . (k (*)) is the operator the solver uses
    (conditioning_mechanisms.py:641-675) and is what forward() applies here; the reference synthesises y with
    a bicubic `Resizer` instead (measurements.py:105), a once-per-image step listed under SURVEY.md 8(f) N3."""

    def __init__(self, in_shape, scale_factor=4, sigma_s=0.1, device=None, kernel=None, kernel_dir=None, **kwargs):
        self._setup(in_shape, sigma_s, device)
        self.scale_factor = scale_factor
        if kernel is None:
            import scipy.io
            kernels = scipy.io.loadmat(_kernel_file("kernels_bicubicx234.mat", kernel_dir))["kernels"]
            k_index = scale_factor - 2 if scale_factor < 5 else 2
            kernel = kernels[0, k_index].astype(np.float64)
        self._make(kernel, sf=scale_factor)
        out = tuple(int(s / scale_factor) for s in self.in_shape[-2:])
        self.out_shape = (1, self.in_shape[-3], *out)


class MaskGenerator:
    """measurements.py:248-320.  Draws from the numpy global RNG in the same order as the reference so the
    masks are bit-identical under the same seed."""

    def __init__(self, mask_type, mask_len_range=None, mask_prob_range=None, image_size=256, margin=(16, 16)):
        assert mask_type in ["box", "random", "both", "extreme"]
        self.mask_type, self.mask_len_range, self.mask_prob_range = mask_type, mask_len_range, mask_prob_range
        self.image_size, self.margin = image_size, margin

    def _retrieve_box(self, shape):
        l, h = int(self.mask_len_range[0]), int(self.mask_len_range[1])
        mask_h = np.random.randint(l, h)
        mask_w = np.random.randint(l, h)
        maxt = self.image_size - self.margin[0] - mask_h
        maxl = self.image_size - self.margin[1] - mask_w
        t = np.random.randint(self.margin[0], maxt)
        le = np.random.randint(self.margin[1], maxl)
        mask = torch.ones(shape)
        mask[..., t:t + mask_h, le:le + mask_w] = 0
        return mask

    def _retrieve_random(self, shape):
        total = self.image_size ** 2
        l, h = self.mask_prob_range
        prob = np.random.uniform(l, h)
        mask_vec = torch.ones([1, total])
        samples = np.random.choice(total, int(total * prob), replace=False)
        mask_vec[:, samples] = 0
        mask_b = mask_vec.view(1, self.image_size, self.image_size).repeat(shape[-3], 1, 1)
        mask = torch.ones(shape)
        mask[:, ...] = mask_b
        return mask

    def __call__(self, img):
        shape = tuple(img.shape)
        if self.mask_type == "random":
            return self._retrieve_random(shape).to(img.device)
        if self.mask_type == "box":
            return self._retrieve_box(shape).to(img.device)
        if self.mask_type == "extreme":
            return (1.0 - self._retrieve_box(shape)).to(img.device)
        raise NotImplementedError(self.mask_type)     # 'both' returns None in the reference as well


@register_operator(name="inpainting")
class InpaintingOperator(_LinearOperator):
    """measurements.py:204-246."""

    def __init__(self, device=None, sigma_s=0.1, mask_opt=None, mask=None, **kwargs):
        size = mask_opt["image_size"] if mask_opt is not None else mask.shape[-1]
        self._setup((1, 3, size, size), sigma_s, device)
        if mask is None:
            torch.randn(*self.in_shape)        # the reference draws an unused image first (measurements.py:244)
            mask = MaskGenerator(**mask_opt)(torch.zeros(*self.in_shape))
        self.mask = torch.as_tensor(mask, dtype=torch.float32).to(self.device).reshape(-1, 3, size, size)
        self.op = engine.Operator.inpaint(self.ctx, self.mask[:, 0])
        self.out_shape = self.in_shape

    def forward(self, data, flatten=False, noiseless=False):
        y = data.clone()
        if not noiseless:
            y = y + self.sigma_s.to(y.dtype) * torch.randn_like(y)
        y = self.A(y).to(data.dtype)
        if flatten:
            idx = torch.where(self.mask > 0)
            return y, y[..., idx[-3], idx[-2], idx[-1]]
        return y

    def transpose(self, data, flatten=False):
        return data.clone()

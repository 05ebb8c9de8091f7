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


def _cubic(t):
    a = t.abs()
    return torch.where(a <= 1, (1.5 * a - 2.5) * a * a + 1, torch.where(a <= 2, ((-0.5 * a + 2.5) * a - 4) * a + 2, torch.zeros_like(a)))


class Resizer(torch.nn.Module):
    """measurement_utils/resizer.py:8-76 for kernel='cubic' over the last two axes: antialiased bicubic resampling with
    reflected borders.  Each dimension is one dense (n_out, n_in) matrix built once in float64 from the definition
    (tap positions, scaled Keys kernel, normalisation, mirror indices: resizer.py:104-165) and applied as a matrix
    product, instead of the reference's gather + broadcast-multiply + sum per dimension.  Once per image."""

    def __init__(self, in_shape, scale_factor=None, output_shape=None, kernel=None, antialiasing=True):
        super().__init__()
        if kernel not in (None, "cubic"):
            raise ValueError("free_hunch_b200.Resizer provides the 'cubic' kernel (the one measurements.py:91 uses)")
        if scale_factor is None:
            scale_factor = [o / i for o, i in zip(output_shape[-2:], in_shape[-2:])]
        if not isinstance(scale_factor, (list, tuple)):
            scale_factor = [scale_factor, scale_factor]
        self.scales = [float(s) for s in scale_factor[-2:]]
        for name, n, sc in zip(("rh", "rw"), in_shape[-2:], self.scales):
            self.register_buffer(name, self._matrix(int(n), sc, antialiasing))

    @staticmethod
    def _matrix(n_in, scale, antialiasing):
        n_out = int(np.ceil(n_in * scale))
        aa = bool(antialiasing) and scale < 1
        width = 4.0 / scale if aa else 4.0
        u = torch.arange(1, n_out + 1, dtype=torch.float64) - (n_out - n_in * scale) / 2.0
        c = u / scale + 0.5 * (1.0 - 1.0 / scale)
        taps = int(np.ceil(width)) + 2
        idx = torch.floor(c - width / 2.0)[:, None] + torch.arange(taps, dtype=torch.float64)[None, :] - 1
        t = c[:, None] - idx - 1.0
        w = scale * _cubic(scale * t) if aa else _cubic(t)
        tot = w.sum(dim=1, keepdim=True)
        w = (w / torch.where(tot == 0, torch.ones_like(tot), tot)).to(torch.float32)      # float32 weights like the reference
        idx = torch.remainder(idx.to(torch.int64), 2 * n_in)
        idx = torch.where(idx >= n_in, 2 * n_in - 1 - idx, idx)
        R = torch.zeros(n_out, n_in, dtype=torch.float32)
        R.index_put_((torch.arange(n_out)[:, None].expand_as(idx), idx), w, accumulate=True)
        return R

    def forward(self, in_tensor):
        x = in_tensor
        rh, rw = self.rh.to(x.dtype), self.rw.to(x.dtype)
        return torch.matmul(torch.matmul(rh, x), rw.transpose(0, 1))


@register_operator(name="super_resolution")
class SuperResolutionOperator(_ConvOperator):
    """measurements.py:87-123.  A = downsample_sf . (k (*)) is the operator the solver uses
    (conditioning_mechanisms.py:641-675, `A` / `AT` / `transpose` here); `forward` synthesises the measurement with the
    antialiased bicubic `Resizer` exactly like the reference (measurements.py:91,105) -- a different map from A."""

    def forward(self, data, flatten=False, noiseless=False):
        y = self.down_sample(data)
        if not noiseless:
            y = y + self.sigma_s.to(y.dtype) * torch.randn_like(y)
        FB, FBC, F2B = self._otf(y)
        self.pre_calculated = (FB, FBC, F2B, None)
        if flatten:
            return y, y.reshape(y.shape[0], -1)
        return y

    def __init__(self, in_shape, scale_factor=4, sigma_s=0.1, device=None, kernel=None, kernel_dir=None, **kwargs):
        self._setup(in_shape, sigma_s, device)
        self.scale_factor = scale_factor
        if kernel is None:
            import scipy.io
            kernels = scipy.io.loadmat(_kernel_file("kernels_bicubicx234.mat", kernel_dir))["kernels"]
            k_index = scale_factor - 2 if scale_factor < 5 else 2
            kernel = kernels[0, k_index].astype(np.float64)
        self._make(kernel, sf=scale_factor)
        self.down_sample = Resizer(self.in_shape, 1 / scale_factor).to(self.device)
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
        if mask is None:
            torch.randn(1, 3, size, size)      # the reference draws an unused image first (measurements.py:244)
            mask = MaskGenerator(**mask_opt)(torch.zeros(1, 3, size, size))
        mask = torch.as_tensor(mask, dtype=torch.float32).reshape(-1, 3, size, size)
        # one mask per image, shared by its channels (measurements.py:287-299 builds it that way); the library applies
        # channel 0 to all three, so a per-channel mask must be refused rather than silently misapplied
        if not (torch.equal(mask[:, 0], mask[:, 1]) and torch.equal(mask[:, 0], mask[:, 2])):
            raise ValueError("InpaintingOperator: the mask must be identical across channels")
        self._setup((mask.shape[0], 3, size, size), sigma_s, device)       # workspace sized for the mask's own batch
        self.mask = mask.to(self.device)
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

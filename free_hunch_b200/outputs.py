"""Output side of a sampling run (reference training/encoders.py:62-73 and
generate_conditional.py:283-287, 499-583): decode to uint8, the reference's file naming, PSNR on the decoded images and
the cross-rank aggregation of per-rank metric sums.  SSIM and LPIPS come from scikit-image / lpips in the reference;
neither package is part of this path (nor of this image), so only PSNR is provided."""
import math

import torch
import torch.distributed as dist


def decode_uint8(x):
    """StandardRGBEncoder.decode (encoders.py:72-73): [-1, 1] floats -> uint8 pixels, (x * 127.5 + 128) clipped and
    TRUNCATED like the reference's `.to(torch.uint8)`."""
    return (x.to(torch.float32) * 127.5 + 128).clip(0, 255).to(torch.uint8)


def image_filename(index, seed):
    """`{idx:06d}_{seed:06d}.png` (generate_conditional.py:283-287)."""
    return f"{int(index):06d}_{int(seed):06d}.png"


def psnr_uint8(a, b):
    """peak_signal_noise_ratio(a, b, data_range=255) of scikit-image (generate_conditional.py:545) for uint8 images:
    10 log10(255^2 / mean((a - b)^2)) in float64; +inf for identical images."""
    d = a.to(torch.float64) - b.to(torch.float64)
    mse = float((d * d).mean())
    return math.inf if mse == 0.0 else 10.0 * math.log10(255.0 ** 2 / mse)


def aggregate_mean(per_rank_value, device=None):
    """generate_conditional.py:557-573: every rank holds the mean of a metric over ITS images; the reported number is the
    all-reduced sum divided by the world size (float32, like the reference).  Without a process group: the value itself."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(per_rank_value)
    t = torch.tensor([per_rank_value], dtype=torch.float32, device=device)
    dist.all_reduce(t)
    return float(t.item()) / dist.get_world_size()

"""Output side of a sampling run (reference training/encoders.py:62-73 and
generate_conditional.py:283-287, 499-583): decode to uint8, the reference's file naming,
PNG files, PSNR and SSIM on the decoded images, the cross-rank aggregation of per-rank metric means, results.txt and the
pruning of excess images.  PSNR and SSIM come from scikit-image in the reference; the package is not in this image, so
both are restated here from its published algorithm (`structural_similarity` defaults: 7 x 7 uniform window, sample
covariance, K1 = 0.01, K2 = 0.03, borders cropped, mean over channels) -- unpinned against scikit-image itself, checked
in tests/test_outputs.py against an independent scipy.ndimage restatement.  LPIPS needs the VGG weights of the `lpips`
package (no network here): `write_results` takes it as an optional number."""
import math
import os

import torch
import torch.distributed as dist
import torch.nn.functional as F


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


def save_images(images, image_paths):
    """generate_conditional.py:283-287: (N, 3, H, W) uint8 -> one RGB PNG per image, directories created on the way."""
    import PIL.Image
    arr = images.permute(0, 2, 3, 1).contiguous().cpu().numpy()
    for a, path in zip(arr, image_paths):
        d = os.path.dirname(path)
        if d:
            os.makedirs(d, exist_ok=True)
        PIL.Image.fromarray(a, "RGB").save(path)


def ssim_uint8(a, b, win_size=7, k1=0.01, k2=0.03):
    """skimage.metrics.structural_similarity(a, b, data_range=255, channel_axis=0) (generate_conditional.py:546) for
    (C, H, W) uint8 images: per channel, local means / sample (co)variances over a win_size x win_size uniform window in
    float64, S = (2 ux uy + C1)(2 vxy + C2) / ((ux^2 + uy^2 + C1)(vx + vy + C2)), mean of S over the pixels whose window
    lies inside the image, then the mean over channels."""
    x = a.to(torch.float64)[:, None]
    y = b.to(torch.float64)[:, None]
    npix = win_size * win_size
    cov_norm = npix / (npix - 1.0)
    mean = lambda t: F.avg_pool2d(t, win_size, stride=1)          # valid windows only == skimage's crop of (win_size - 1) // 2
    ux, uy = mean(x), mean(y)
    vx = cov_norm * (mean(x * x) - ux * ux)
    vy = cov_norm * (mean(y * y) - uy * uy)
    vxy = cov_norm * (mean(x * y) - ux * uy)
    c1, c2 = (k1 * 255.0) ** 2, (k2 * 255.0) ** 2
    s_map = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))
    return float(s_map.flatten(1).mean(dim=1).mean())


def batch_metrics(cond_images, gen_images):
    """The per-batch metric loop of generate_conditional.py:543-547: sums of PSNR and SSIM over the (conditioning image,
    generated image) pairs of one batch (uint8, (N, 3, H, W))."""
    psnr = ssim = 0.0
    for c, g in zip(cond_images, gen_images):
        psnr += psnr_uint8(c, g)
        ssim += ssim_uint8(c, g)
    return psnr, ssim


def write_results(outdir, psnr, ssim, lpips=float("nan")):
    """results.txt exactly as generate_conditional.py:575-579 writes it."""
    os.makedirs(outdir, exist_ok=True)
    path = os.path.join(outdir, "results.txt")
    with open(path, "w") as f:
        f.write(f"PSNR: {psnr}\n")
        f.write(f"SSIM: {ssim}\n")
        f.write(f"LPIPS: {lpips}\n")
    return path


def prune_images(outdir, keep, subdirs=("images", "cond_images", "forward_images")):
    """generate_conditional.py:581-590 (rank 0): keep the first `keep` PNGs (sorted by name) of every image directory."""
    removed = 0
    for sd in subdirs:
        d = os.path.join(outdir, sd)
        if not os.path.isdir(d):
            continue
        for name in sorted(f for f in os.listdir(d) if f.endswith(".png"))[keep:]:
            os.remove(os.path.join(d, name))
            removed += 1
    return removed

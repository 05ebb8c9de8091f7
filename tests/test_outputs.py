"""Output side (generate_conditional.py:283-287, 499-590) and the dct_variance extraction (N3:
do_frequency_analysis.py:26-72) -- host-side logic, CPU."""
import os

import numpy as np
import pytest
import scipy.ndimage as ndi
import torch

import helpers as H
from free_hunch_b200 import outputs as O


def _ssim_scipy(a, b, win=7):
    """Independent restatement of skimage.metrics.structural_similarity (uniform_filter, reflect borders, crop)."""
    vals = []
    for x, y in zip(a.astype(np.float64), b.astype(np.float64)):
        f = lambda t: ndi.uniform_filter(t, size=win)
        ux, uy = f(x), f(y)
        n = win * win
        cn = n / (n - 1)
        vx, vy, vxy = cn * (f(x * x) - ux * ux), cn * (f(y * y) - uy * uy), cn * (f(x * y) - ux * uy)
        c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
        s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
        p = (win - 1) // 2
        vals.append(s[p:-p, p:-p].mean())
    return float(np.mean(vals))


def test_ssim_psnr_against_independent_restatement():
    rs = np.random.RandomState(0)
    base = ndi.gaussian_filter(rs.uniform(0, 255, (3, 64, 80)), (0, 2, 2))
    a = np.clip(base, 0, 255).astype(np.uint8)
    b = np.clip(base + rs.normal(0, 12, base.shape), 0, 255).astype(np.uint8)
    ta, tb = torch.from_numpy(a), torch.from_numpy(b)
    assert abs(O.ssim_uint8(ta, tb) - _ssim_scipy(a, b)) < 1e-12
    assert O.ssim_uint8(ta, ta) == pytest.approx(1.0, abs=1e-15)
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    assert O.psnr_uint8(ta, tb) == pytest.approx(10 * np.log10(255 ** 2 / mse), rel=1e-14)
    p, s = O.batch_metrics(torch.stack([ta, tb]), torch.stack([tb, tb]))
    assert s == pytest.approx(O.ssim_uint8(ta, tb) + 1.0) and p == float("inf")


def test_png_results_and_pruning(tmp_path):
    import PIL.Image
    rs = np.random.RandomState(1)
    imgs = torch.from_numpy(rs.randint(0, 256, (3, 3, 16, 12)).astype(np.uint8))
    out = str(tmp_path)
    paths = [os.path.join(out, "images", O.image_filename(i, 16)) for i in range(3)]
    O.save_images(imgs, paths)
    O.save_images(imgs[:2], [os.path.join(out, "cond_images", O.image_filename(i, 16)) for i in range(2)])
    back = np.asarray(PIL.Image.open(paths[1]))                         # (H, W, 3), lossless
    assert back.shape == (16, 12, 3) and np.array_equal(back, imgs[1].permute(1, 2, 0).numpy())
    assert os.path.basename(paths[2]) == "000002_000016.png"
    res = O.write_results(out, 27.5, 0.81, 0.2)
    assert open(res).read() == "PSNR: 27.5\nSSIM: 0.81\nLPIPS: 0.2\n"
    assert O.prune_images(out, 1) == 3
    assert sorted(os.listdir(os.path.join(out, "images"))) == ["000000_000016.png"]
    assert sorted(os.listdir(os.path.join(out, "cond_images"))) == ["000000_000016.png"]


@pytest.mark.skipif(H.HAS_CUDA, reason="CPU emulation is for boxes without a GPU")
def test_dct_variance_accumulator_emulated():
    """DCTVarianceAccumulator over the library's own DCT (CPU emulation build) against scipy's orthonormal DCT-II, the
    `limit` clipping of do_frequency_analysis.py:36-39 and the file CovarianceHessianBFGSDCT loads."""
    import scipy.fft as sf
    from free_hunch_b200 import covariance as C
    E = H.use_emulation()
    ctx = E.Context("cpu", max_batch=2)
    rs = np.random.RandomState(2)
    x = np.clip(ndi.gaussian_filter(rs.standard_normal((5, 3, 256, 256)), (0, 0, 3, 3)) * 4, -1, 1).astype(np.float32)
    acc = C.DCTVarianceAccumulator(ctx)
    for i in range(0, 5, 2):
        n = acc.update(torch.from_numpy(x[i:i + 2]), limit=4)
    assert n == 4 and acc.count == 4
    c = sf.dctn(x[:4].astype(np.float64), axes=(-2, -1), norm="ortho")
    ref = (c ** 2).mean(0) - c.mean(0) ** 2
    got = acc.variance().numpy()
    assert got.dtype == np.float32 and got.shape == (3, 256, 256)
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-5

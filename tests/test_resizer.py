"""the bicubic `Resizer` the reference uses to synthesise super-resolution measurements
(measurement_utils/resizer.py, measurements.py:91,105).  Golden vectors come from the reference's own class
(tests/golden/make_golden_resizer.py); checked are the float64 oracle restatement and the product's matrix form, on
the 256 -> 64 / 128 case and on ragged sizes (30 x 21, factors 4 and 3: reflected borders, non-integer centres)."""
import numpy as np
import pytest
import torch

import helpers as H
from helpers import rel_l2

from oracle import fh_oracle as fo  # noqa: E402

G = np.load(H.ROOT + "/tests/golden/resizer.npz")


@pytest.mark.parametrize("x,y,sf", [("x256", "y256_sf4", 4), ("x256", "y256_sf2", 2), ("x_ragged", "y_ragged_sf4", 4), ("x_ragged", "y_ragged_sf3", 3)])
def test_oracle_resizer_vs_reference(x, y, sf):
    out = fo.bicubic_downsample(G[x], sf)
    assert out.shape == G[y].shape
    assert rel_l2(out, G[y]) < 1e-6


@pytest.mark.parametrize("x,y,sf", [("x256", "y256_sf4", 4), ("x256", "y256_sf2", 2), ("x_ragged", "y_ragged_sf4", 4), ("x_ragged", "y_ragged_sf3", 3)])
def test_product_resizer_vs_reference(x, y, sf):
    from free_hunch_b200.measurements import Resizer
    xin = torch.from_numpy(G[x])
    out = Resizer(tuple(xin.shape), 1 / sf)(xin)
    assert tuple(out.shape) == G[y].shape and out.dtype == torch.float32
    assert rel_l2(out, G[y]) < 1e-6
    out64 = Resizer(tuple(xin.shape), 1 / sf)(xin.double())            # the sampler carries float64 images
    assert out64.dtype == torch.float64 and rel_l2(out64, fo.bicubic_downsample(G[x], sf)) < 1e-7


def test_resizer_matrix_properties():
    R = fo.resizer_matrix(256, 0.25)
    assert R.shape == (64, 256)
    assert np.allclose(R.sum(axis=1), 1.0, atol=1e-6)                   # normalised rows: constants are preserved
    assert np.allclose(R[10, 32:54], R[11, 36:58], atol=1e-7)           # interior rows are shifts of each other by 4
    assert (np.count_nonzero(R, axis=1) <= 18).all()                    # ceil(16) + 2 taps

"""CPU check of the coverage argument of the hash-grid patch gathering (csrc/bx_patches.cu, hg_coord / hg_inv_cell): a point
closer than the radius to a key-point lies in one of the 27 cells around the key-point's cell -- in fp32, with the kernel's own
arithmetic (cell edge 1.001 * radius, coordinate = floor(fl(v * fl(1 / fl(r * 1.001))))), for cloud extents up to the documented
limit of 4096 cells per axis.  The GPU test (test_select_patches_grid_equals_scan) compares whole results; this pins the bound."""
import numpy as np
import pytest


def _cell(v, r):
    ic = np.float32(1.0) / (np.float32(r) * np.float32(1.001))
    return np.floor((v.astype(np.float32) * ic).astype(np.float32)).astype(np.int64)


@pytest.mark.parametrize("radius,extent", [(0.05, 3.0), (0.3, 60.0), (1.5, 120.0), (0.0123, 50.0), (2.0, 8000.0)])
def test_neighbour_within_radius_is_within_one_cell(radius, extent):
    rng = np.random.default_rng(int(radius * 1e4) + int(extent))
    n = 400000
    q = rng.uniform(-extent, extent, size=(n, 3)).astype(np.float32)
    # offsets of length just below the radius in random directions, plus axis-aligned worst cases
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[: n // 4] = np.eye(3)[rng.integers(0, 3, n // 4)] * rng.choice([-1.0, 1.0], size=(n // 4, 1))
    scale = rng.uniform(0.9, 0.9999999, size=(n, 1)) * radius
    p = (q.astype(np.float64) + d * scale).astype(np.float32)
    # keep the pairs that the kernel's own fp32 test accepts: d2 = ((dx*dx)+(dy*dy))+(dz*dz) < r*r
    dx, dy, dz = (q[:, 0] - p[:, 0]), (q[:, 1] - p[:, 1]), (q[:, 2] - p[:, 2])
    d2 = ((dx * dx).astype(np.float32) + (dy * dy).astype(np.float32)).astype(np.float32) + (dz * dz).astype(np.float32)
    hit = d2.astype(np.float32) < np.float32(radius) * np.float32(radius)
    assert hit.mean() > 0.5
    assert extent / (radius * 1.001) <= 4096.0                     # the documented limit; the last case sits at ~3996 cells
    dc = np.abs(_cell(q, radius) - _cell(p, radius))
    assert dc[hit].max() <= 1, f"a hit {dc[hit].max()} cells away from its key-point"


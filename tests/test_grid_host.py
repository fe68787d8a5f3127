"""Host geometry (medicalsemseg_b200.grid + the C library's host-only helpers) against the oracle."""
import numpy as np
import pytest

from medicalsemseg_b200 import grid as G
from medicalsemseg_b200 import importance as I
from oracle import monai08 as M
from oracle import sliding_window as osw


def cases():
    rs = np.random.RandomState(0)
    out = [((128, 128, 128), 96, 0.25), ((512, 512, 200), (96, 96, 96), 0.5), ((512, 512, 1024), 96, 0.5),
           ((240, 240, 155), 96, 0.5), ((10, 40, 21), 16, 0.25), ((16, 40, 40), (16, -1, 24), 0.25)]
    for _ in range(200):
        shape = tuple(int(v) for v in rs.randint(1, 70, size=3))
        roi = tuple(int(v) for v in rs.randint(1, 40, size=3))
        out.append((shape, roi, float(rs.choice([0.0, 0.1, 0.25, 0.5, 0.6, 0.75, 0.9, 0.99]))))
    return out


def test_grid_matches_oracle_everywhere():
    for shape, roi, ov in cases():
        g = G.make_grid(shape, roi, ov)
        roi_t = M.fall_back_tuple(roi, shape)
        image = tuple(max(a, b) for a, b in zip(shape, roi_t))
        interval, starts = osw.window_grid(image, roi_t, ov)
        assert g.roi == roi_t and g.image_size == image and g.interval == interval
        assert [list(s) for s in g.starts] == starts, (shape, roi, ov)
        slices = M.dense_patch_slices(image, roi_t, interval)
        assert g.n_windows == len(slices)
        for n in (0, g.n_windows // 2, g.n_windows - 1):
            assert g.window_start(n) == tuple(s.start for s in slices[n])
            stops = [s.stop for s in slices[n]]
            assert list(g.centers(n)) == osw.window_centers(stops, roi_t, image)


def test_cover_table_is_exact():
    for shape, roi, ov in cases()[:60]:
        g = G.make_grid(shape, roi, ov)
        t = g.table
        assert t[0] == 0x4D535331
        for a in range(3):
            st = t[t[10 + a]: t[10 + a] + t[7 + a]]
            assert list(st) == list(g.starts[a])
            cov = t[t[13 + a]: t[13 + a] + g.image_size[a]]
            for x in range(g.image_size[a]):
                want = [i for i, s in enumerate(g.starts[a]) if s <= x < s + g.roi[a]]
                lo, hi = int(cov[x]) & 0xFFFF, int(cov[x]) >> 16
                assert list(range(lo, hi)) == want


def test_reference_error_behaviour():
    with pytest.raises(AssertionError, match="overlap must be >= 0 and < 1."):
        G.make_grid((32, 32, 32), 16, 1.0)
    with pytest.raises(AssertionError):
        G.make_grid((32, 32, 32), 16, -0.1)
    with pytest.raises(ValueError):
        G.make_grid((32, 32, 32), (16, 16), 0.5)
    with pytest.raises(ValueError):
        G._option("nearest", G.PAD_MODES, "padding_mode")
    assert G._option(M.PytorchPadMode.REFLECT, G.PAD_MODES, "padding_mode") == "reflect"


@pytest.mark.parametrize("roi", [(96, 96, 96), (16, 16, 16), (24, 16, 32), (8, 12, 20), (5, 7, 9)])
def test_host_profiles_rebuild_the_monai08_map(roi):
    import torch
    want = M.compute_importance_map(roi, mode="gaussian", sigma_scale=0.125)
    p = [I.monai08_profile(n, n * 0.125) for n in roi]
    outer = (p[0][:, None, None] * p[1][None, :, None]) * p[2][None, None, :]
    vmax = (p[0].max() * p[1].max()) * p[2].max()
    got = outer / vmax
    got = torch.clamp(got, min=got[got != 0].min().item())
    assert torch.equal(got, want)


def test_host_profiles_rebuild_the_monai12_map():
    import torch
    roi = (24, 16, 32)
    want = M.compute_importance_map_v12(roi)
    p = [I.monai12_profile(n, n * 0.125) for n in roi]
    outer = (p[0][:, None, None] * p[1][None, :, None]) * p[2][None, None, :]
    assert torch.equal(torch.clamp(outer, min=max(outer.min().item(), 1e-3)), want)

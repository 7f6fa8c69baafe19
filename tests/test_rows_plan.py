"""CPU: the host arithmetic of the default-mode resize kernel (csrc/k_resize_rows.cu) through fm_debug_rows_plan -- no
device needed.  The library re-checks its own tables tap by tap (every read inside its TMA box, every weight the one
cv2's INTER_AREA table gives); here the plan's shape is checked against the oracle's tables and the kernel's limits over a
sweep of geometries, including the ones the kernel must leave to the warp-per-row kernels."""
import ctypes as C

import pytest

from oracle import restated as R


@pytest.fixture(scope="module")
def lib():
    from find_motion_b200 import build, _lib
    build.build()
    return _lib.load()


def _plan(lib, W, H, box):
    from find_motion_b200 import _lib
    info = _lib.fm_rows_plan_info()
    _lib.check(lib.fm_debug_rows_plan(W, H, box, C.byref(info)))
    return {f: getattr(info, f) for f, _ in info._fields_}


SIZES = [(1920, 1080), (3840, 2160), (1280, 720), (640, 480), (2560, 1440), (1936, 1096), (4096, 2160), (800, 600),
         (636, 476), (1918, 1080)]                      # the last two: rows that are not a multiple of 16 bytes
BOXES = [32, 64, 100, 111, 123, 150, 200, 300, 333, 500, 640, 900]


@pytest.mark.parametrize("W,H", SIZES)
def test_plan_invariants(lib, W, H):
    for box in BOXES:
        if box > W:
            continue
        p = _plan(lib, W, H, box)                       # raises if the library's own tap-by-tap check fails
        w, h = box, int(H * (box / float(W)))
        if h < 1:
            continue
        xt, yt = R.area_tab(W, w), R.area_tab(H, h)
        taps = max(len(t) for t in xt)
        if not p["usable"]:
            sx, sy = W / w, H / h
            integer = abs(sx - round(sx)) < 1e-12 and abs(sy - round(sy)) < 1e-12
            assert (W * 3) % 16 != 0 or integer or taps < 4 or w == W, (W, H, box, p)
            continue
        assert (W * 3) % 16 == 0
        assert 1 <= p["box_rows"] <= 160 and p["band_rows"] >= 1
        assert p["box_bytes"] % 16 == 0 and (p["box_bytes"] // 16) % 2 == 1 and p["box_bytes"] <= 1024
        assert p["smem_bytes"] <= 200 * 1024
        assert p["chunk_cols"] in (1, 2) and p["seg_cols"] % p["chunk_cols"] == 0
        assert p["bands"] == -(-h // p["band_rows"]) and p["segs"] == -(-w // p["seg_cols"])
        if p["chunk_cols"] == 2:
            assert p["smem_bytes"] <= 56 * 1024              # two columns per box only while four CTAs fit an SM
        assert p["max_groups"] == -(-taps // 4)
        rows = max(yt[min(d0 + p["band_rows"], h) - 1][-1][0] - yt[d0][0][0] + 1 for d0 in range(0, h, p["band_rows"]))
        assert rows == p["box_rows"]                       # the tallest band's source rows are the box


def test_reference_default_geometry(lib):
    """1080p -> box 100 (the reference's CLI default): the configuration bench.py times."""
    p = _plan(lib, 1920, 1080, 100)
    assert p == dict(usable=1, band_rows=8, box_rows=156, chunk_cols=2, seg_cols=20, box_bytes=144, smem_bytes=p["smem_bytes"],
                     bands=7, segs=5, max_groups=5)
    assert 4 * (p["smem_bytes"] + 1024) <= 228 * 1024        # four CTAs per SM


def test_bad_arguments_are_loud(lib):
    from find_motion_b200 import _lib
    info = _lib.fm_rows_plan_info()
    with pytest.raises(_lib.FmError):
        _lib.check(lib.fm_debug_rows_plan(640, 480, 800, C.byref(info)))     # upscaling

"""CPU: stage-by-stage check of the restated oracle against the cv2 build that pins parity
(opencv-python-headless 4.13), and against the real reference loop when
/root/reference is mounted (build container only).  Skipped where cv2 is absent."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import ref_loader, restated as R  # noqa: E402
from find_motion_b200 import synth  # noqa: E402


def test_gray():
    rng = np.random.default_rng(0)
    x = rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)
    assert (R.bgr2gray(x) == cv2.cvtColor(x, cv2.COLOR_BGR2GRAY)).all()


@pytest.mark.parametrize("k", [1, 3, 5, 7, 9, 11, 15, 17, 21, 33, 49, 65, 97, 193])
def test_gaussian_blur(k):
    rng = np.random.default_rng(k)
    for shp in [(75, 100), (240, 320), (40, 30)]:
        g = rng.integers(0, 256, shp, dtype=np.uint8)
        assert (R.gaussian_blur(g, k) == cv2.GaussianBlur(g, (k, k), 0)).all(), (k, shp)


@pytest.mark.parametrize("W,H,w", [(640, 480, 100), (1920, 1080, 100), (1280, 720, 100), (1920, 1080, 300),
                                   (640, 480, 320), (640, 480, 160), (600, 480, 200), (640, 480, 640),
                                   (333, 217, 100), (3840, 2160, 100)])
def test_resize_area(W, H, w):
    rng = np.random.default_rng(W + w)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    h = int(H * (w / float(W)))
    assert (R.resize_area(img, w, h) == cv2.resize(img, (w, h), interpolation=cv2.INTER_AREA)).all()


def test_dilate():
    rng = np.random.default_rng(1)
    t = (rng.random((75, 100)) > 0.97).astype(np.uint8) * 255
    assert (R.dilate5(t) == cv2.dilate(t, None, iterations=2)).all()


@pytest.mark.parametrize("shape", [(75, 100), (56, 100), (33, 47), (240, 320)])
def test_background(shape):
    rng = np.random.default_rng(shape[0])
    for alpha in [0.1, 0.05, 0.3, 0.01, 0.5]:
        bg = rng.random(shape) * 255
        ref = bg.copy()
        for _ in range(4):
            src = rng.integers(0, 256, shape, dtype=np.uint8)
            cv2.accumulateWeighted(src, ref, alpha)
            bg = R.accumulate_weighted(bg, src, alpha)
            assert (bg == ref).all(), (shape, alpha)
            assert (R.bg_to_u8(bg) == cv2.convertScaleAbs(ref)).all()


def test_bg_to_u8_ties():
    base = np.arange(0, 255, dtype=np.float64) + 0.5
    offs = np.array([0.0, 1e-9, -1e-9, 1e-6, -1e-6, 2.0 ** -30, -2.0 ** -30])
    x = (base[:, None] + offs[None, :]).copy()
    assert (R.bg_to_u8(x) == cv2.convertScaleAbs(x)).all()


def test_masks_polygons():
    rng = np.random.default_rng(3)
    for _ in range(300):
        w, h = int(rng.integers(20, 200)), int(rng.integers(20, 150))
        n = int(rng.integers(3, 7))
        ang = np.sort(rng.random(n) * 2 * np.pi)
        if rng.random() < 0.5:
            ang = ang[::-1]
        cx, cy = rng.integers(-10, w + 10), rng.integers(-10, h + 10)
        rad = rng.integers(3, max(w, h))
        pts = [(int(cx + rad * np.cos(a)), int(cy + rad * np.sin(a))) for a in ang]
        if rng.random() < 0.2:
            pts = [(int(rng.integers(-20, w + 20)), int(rng.integers(-20, h + 20))) for _ in range(n)]
        img = np.full((h, w), 255, np.uint8)
        cv2.fillConvexPoly(img, np.array(pts, np.int32), 0)
        m = np.zeros((h, w), bool)
        R.fill_convex_poly(m, pts)
        assert ((img == 0) == m).all(), pts


def test_masks_rectangles_and_readme():
    rng = np.random.default_rng(4)
    for _ in range(100):
        w, h = int(rng.integers(20, 200)), int(rng.integers(20, 150))
        a = [(int(rng.integers(-20, w + 20)), int(rng.integers(-20, h + 20))) for _ in range(2)]
        img = np.full((h, w), 255, np.uint8)
        cv2.rectangle(img, a[0], a[1], 0, cv2.FILLED)
        assert ((img == 0) == R.rasterise_masks(w, h, [a], 1.0)).all()
    for (W, H, box) in [(1920, 1080, 1920), (1920, 1080, 100), (640, 480, 100), (1280, 720, 1280)]:
        scale = box / W
        h = int(H * (box / float(W)))
        img = np.full((h, box), 255, np.uint8)
        for area in synth.CFG2_MASKS:
            pts = R.scale_area(area, scale)
            if len(pts) == 2:
                cv2.rectangle(img, pts[0], pts[1], 0, cv2.FILLED)
            else:
                cv2.fillConvexPoly(img, np.array(pts, np.int32), 0)
        assert ((img == 0) == R.rasterise_masks(box, h, synth.CFG2_MASKS, scale)).all()


def test_external_components():
    rng = np.random.default_rng(5)
    for _ in range(300):
        h, w = int(rng.integers(8, 60)), int(rng.integers(8, 80))
        dens = rng.choice([0.02, 0.1, 0.3, 0.5, 0.7])
        t = (rng.random((h, w)) < dens).astype(np.uint8) * 255
        if rng.random() < 0.5:
            t = cv2.dilate(t, None, iterations=int(rng.integers(1, 3)))
        if rng.random() < 0.5:
            t[rng.random((h, w)) < 0.05] = 0
        cnts = cv2.findContours(t.copy(), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[-2]
        ref = sorted((int(round(2 * cv2.contourArea(c))), tuple(int(v) for v in cv2.boundingRect(c)))
                     for c in cnts)
        assert R.external_components(t) == ref


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference not mounted")
def test_against_live_reference_loop():
    """Drive the real VideoMotion loop and the oracle over a fresh clip (not a committed golden)."""
    from tests import helpers

    frames = synth.make_clip(400, 300, 60, seed=99, fps=10)
    kw = dict(fps=10, box_size=200, blur_scale=15, threshold=9, avg=0.2, min_time=0.3, cache_time=0.7,
              min_box_scale=50, mask_areas=[((0, 0), (100, 100)), ((0, 0), (0, 100), (100, 0))])
    ref = ref_loader.run_reference(list(frames), **kw)
    so = R.StreamOracle(400, 300, **kw)
    for t, gold in enumerate(ref["trace"]):
        rec = so.process(frames[t], keep_planes=True)
        helpers.check_record("live", t, rec, gold, rec["planes"])

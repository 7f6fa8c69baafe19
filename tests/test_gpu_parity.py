"""GPU parity: the CUDA path (through the C ABI, libfmgpu.so) against the golden traces recorded
from the real reference and against the CPU oracle on seeded inputs.  Bit-exact everywhere:
gray, blur, dilated threshold, float64 background, contour areas/boxes, decisions."""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu


def _engine(fx, T, **extra):
    from find_motion_b200.engine import MotionEngine
    c = fx["clip"]
    return MotionEngine(c["W"], c["H"], n_streams=1, max_frames=T, keep_planes=True, **fx["kwargs"], **extra)


def _run_and_check(name, fx, clip, T, eng):
    import torch
    n = clip.shape[0]
    dev = torch.from_numpy(clip).cuda()
    for key in ("gaussian", "min_area", "max_area", "cache_frames", "min_movement_frames", "scale"):
        assert eng.info[key] == fx["params"][key], key
    for t0 in range(0, n, T):
        t1 = min(n, t0 + T)
        stats = eng.process(dev[None, t0:t1])
        for t in range(t0, t1):
            gold = fx["trace"][t]
            st = stats[0, t - t0]
            pl = eng.planes(0, t - t0, bg=(t == t1 - 1))
            ncomp, comps = eng.components(0, t - t0)
            assert ncomp == len(gold["areas"]), f"{name} frame {t}: contour count {ncomp} vs {len(gold['areas'])}"
            rec = {"areas": sorted(a / 2.0 for a, _ in comps), "boxes": sorted(b for _, b in comps),
                   "movement": bool(st["movement"]), "counter": int(st["movement_counter"]),
                   "decay": int(st["movement_decay"]), "cache_len": int(st["cache_len"]),
                   "wrote": bool(st["wrote"]), "n_flush": int(st["n_flush"])}
            helpers.check_record(name, t, rec, gold, pl)
            assert int(st["n_contours"]) == len(gold["areas"])


@pytest.mark.parametrize("T", [1, 7])
@pytest.mark.parametrize("name", helpers.golden_names())
def test_golden_traces(name, T):
    fx = helpers.load_golden(name)
    clip = helpers.golden_clip(fx)
    with _engine(fx, T) as eng:
        _run_and_check(name, fx, clip, T, eng)


def test_reset_restarts_stream():
    fx = helpers.load_golden("full_256x192_k5")
    clip = helpers.golden_clip(fx)[:24]
    with _engine(fx, 8) as eng:
        _run_and_check("first", fx, clip, 8, eng)
        eng.reset()
        _run_and_check("after-reset", fx, clip, 8, eng)


def test_contours_random_planes():
    from find_motion_b200.engine import label_components
    from oracle import restated as R
    rng = np.random.default_rng(11)
    for trial in range(120):
        h, w = int(rng.integers(1, 90)), int(rng.integers(1, 150))
        dens = rng.choice([0.0, 0.02, 0.1, 0.3, 0.5, 0.7, 1.0])
        t = (rng.random((h, w)) < dens).astype(np.uint8) * 255
        if rng.random() < 0.5:
            t = R.dilate5(t)
        if rng.random() < 0.5:
            t[rng.random((h, w)) < 0.05] = 0
        assert label_components(t) == R.external_components(t), (trial, h, w)


def test_contours_random_planes_all_lane_groupings():
    """Widths that select 8, 16 and 32 lanes per row in the labelling kernel (<= 256, <= 512, wider)."""
    from find_motion_b200.engine import label_components
    from oracle import restated as R
    rng = np.random.default_rng(17)
    for trial in range(60):
        w = int(rng.integers(*[(150, 257), (257, 513), (513, 1100)][trial % 3]))
        h = int(rng.integers(2, 70))
        dens = rng.choice([0.01, 0.05, 0.2, 0.5, 0.8])
        t = (rng.random((h, w)) < dens).astype(np.uint8) * 255
        if rng.random() < 0.6:
            t = R.dilate5(t)
        if rng.random() < 0.5:
            t[rng.random((h, w)) < 0.03] = 0
        assert label_components(t) == R.external_components(t), (trial, h, w)


def test_contours_structured_planes():
    """Nested rings, spirals and border-touching shapes (RETR_EXTERNAL nesting / hole filling)."""
    from find_motion_b200.engine import label_components
    from oracle import restated as R
    h, w = 96, 200
    t = np.zeros((h, w), np.uint8)
    for i, r in enumerate(range(4, 44, 4)):          # concentric square rings
        if i % 2 == 0:
            t[48 - r:48 + r, 50 - r:50 + r] = 255
        else:
            t[48 - r:48 + r, 50 - r:50 + r] = 0
    t = np.maximum(t, np.flipud(np.fliplr(t)))
    t[0, :] = 255
    t[:, -1] = 255                                   # frame along two borders
    t[10:80, 120] = 255
    t[10, 120:190] = 255
    t[80, 120:190] = 255                              # open U shape: no hole
    t[30:60, 150:170] = 255
    t[35:55, 155:165] = 0                             # box with hole ...
    t[44:46, 159:161] = 255                           # ... with a dot inside
    assert label_components(t) == R.external_components(t)
    comb = np.zeros((64, 257), np.uint8)
    comb[::2, :] = 255
    comb[:, ::4] = 255
    assert label_components(comb) == R.external_components(comb)
    full = np.full((33, 65), 255, np.uint8)
    assert label_components(full) == R.external_components(full)


def test_contours_heavy_planes_use_the_global_fallback():
    """More runs than the shared-memory run table holds (10240): the global-memory kernel takes over."""
    from find_motion_b200.engine import label_components
    from oracle import restated as R
    rng = np.random.default_rng(13)
    t = (rng.random((300, 1024)) < 0.35).astype(np.uint8) * 255       # ~70k runs
    assert label_components(t, max_n=200000) == R.external_components(t)
    stripes = np.zeros((220, 2048), np.uint8)
    stripes[:, ::2] = 255                                              # 1024 runs per row
    stripes[100:120, :] = 255
    assert label_components(stripes, max_n=200000) == R.external_components(stripes)
    wide = np.zeros((40, 5000), np.uint8)                              # wider than 4096 px: global kernel only
    wide[5:30, 100:4900] = 255
    wide[10:20, 2000:3000] = 0
    assert label_components(wide) == R.external_components(wide)


def test_mask_raster_random_polygons():
    from find_motion_b200.engine import MotionEngine
    from oracle import restated as R
    rng = np.random.default_rng(12)
    for trial in range(40):
        W, H = int(rng.integers(40, 400)), int(rng.integers(40, 300))
        box = int(rng.integers(20, W + 1))
        areas = []
        for _ in range(int(rng.integers(1, 5))):
            if rng.random() < 0.3:
                areas.append(tuple((int(rng.integers(-20, W + 20)), int(rng.integers(-20, H + 20))) for _ in range(2)))
            else:
                n = int(rng.integers(3, 8))
                ang = np.sort(rng.random(n) * 2 * np.pi)
                cx, cy, rad = rng.integers(0, W), rng.integers(0, H), rng.integers(5, max(W, H))
                areas.append(tuple((int(cx + rad * np.cos(a)), int(cy + rad * np.sin(a))) for a in ang))
        with MotionEngine(W, H, box_size=box, mask_areas=areas) as eng:
            got = eng.mask(0).astype(bool)
            want = R.rasterise_masks(eng.w, eng.h, areas, eng.info["scale"])
            assert (got == want).all(), (trial, W, H, box, areas)


def test_streams_are_independent():
    """Three streams batched in one context == the three streams processed alone by the oracle."""
    import torch
    from find_motion_b200 import synth
    from find_motion_b200.engine import MotionEngine
    from oracle import restated as R
    W, H, n, T = 160, 120, 30, 6
    kw = dict(fps=6, box_size=160, blur_scale=32, threshold=10, avg=0.2, min_time=0.5, cache_time=1.0,
              min_box_scale=50, mask_areas=[((5, 5), (40, 30))])
    clips = np.stack([synth.make_clip(W, H, n, seed=50 + s, fps=6) for s in range(3)])
    oracles = [R.StreamOracle(W, H, **kw) for _ in range(3)]
    dev = torch.from_numpy(clips).cuda()
    with MotionEngine(W, H, n_streams=3, max_frames=T, keep_planes=True, **kw) as eng:
        for t0 in range(0, n, T):
            stats = eng.process(dev[:, t0:t0 + T])
            for s in range(3):
                for t in range(t0, t0 + T):
                    rec = oracles[s].process(clips[s, t], keep_planes=True)
                    st = stats[s, t - t0]
                    pl = eng.planes(s, t - t0, gray=True, blur=True, thresh=True, bg=(t == t0 + T - 1))
                    assert (pl["gray"] == rec["planes"]["gray"]).all()
                    assert (pl["blur"] == rec["planes"]["blur"]).all()
                    assert (pl["thresh"] == rec["planes"]["thresh"]).all()
                    if "bg" in pl and t == t0 + T - 1:
                        assert (pl["bg"] == rec["planes"]["bg"]).all()
                    _, comps = eng.components(s, t - t0)
                    assert sorted(a / 2.0 for a, _ in comps) == rec["areas"]
                    assert sorted(b for _, b in comps) == rec["boxes"]
                    for a, b in (("movement", "movement"), ("movement_counter", "counter"),
                                 ("movement_decay", "decay"), ("cache_len", "cache_len"), ("wrote", "wrote"),
                                 ("n_flush", "n_flush")):
                        assert int(st[a]) == int(rec[b]), (s, t, a)


def test_host_entry_point_matches_device_entry_point():
    import torch
    from find_motion_b200 import synth
    from find_motion_b200.engine import MotionEngine
    W, H, n = 192, 108, 12
    kw = dict(fps=6, box_size=96, blur_scale=20, threshold=10)
    clips = np.stack([synth.make_clip(W, H, n, seed=70 + s, fps=6) for s in range(2)])
    with MotionEngine(W, H, n_streams=2, max_frames=n, **kw) as a, \
            MotionEngine(W, H, n_streams=2, max_frames=n, **kw) as b:
        sa = a.process(torch.from_numpy(clips).cuda())
        sb = b.process_host(clips)
        assert (sa == sb).all()


def test_errors_are_loud():
    from find_motion_b200 import _lib
    from find_motion_b200.engine import MotionEngine
    with pytest.raises(_lib.FmError):
        MotionEngine(640, 480, box_size=800)          # upscaling resize is out of scope
    with pytest.raises(_lib.FmError):
        MotionEngine(0, 480)
    import torch
    with MotionEngine(64, 48, max_frames=2, box_size=64) as eng:
        with pytest.raises(_lib.FmError):
            eng.process(torch.zeros((1, 3, 48, 64, 3), dtype=torch.uint8, device="cuda"))

"""GPU parity at BASELINE.json's full sizes: oracle comparisons on a few frames per configuration
plus size-independent properties on the full batch shapes (replicated streams must agree with each
other and with the oracle; reset + replay is idempotent; chunking does not change anything)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_check(W, H, n, T, kw, seed, keep=True, check_planes=("gray", "blur", "thresh")):
    import torch
    from find_motion_b200 import synth
    from find_motion_b200.engine import MotionEngine
    from oracle import restated as R
    clip = synth.make_clip(W, H, n, seed=seed, fps=kw.get("fps", 30),
                           script=[("walker", 1, n), ("ring", 2, n), ("blip", 1, 3)])
    orc = R.StreamOracle(W, H, **kw)
    dev = torch.from_numpy(clip).cuda()
    with MotionEngine(W, H, n_streams=1, max_frames=T, keep_planes=keep, **kw) as eng:
        for t0 in range(0, n, T):
            t1 = min(n, t0 + T)
            stats = eng.process(dev[None, t0:t1])
            for t in range(t0, t1):
                rec = orc.process(clip[t], keep_planes=True)
                pl = eng.planes(0, t - t0, gray=keep, blur=keep, bg=(t == t1 - 1))
                for key in check_planes:
                    if key in pl:
                        assert (pl[key] == rec["planes"][key]).all(), (key, t)
                if t == t1 - 1:
                    assert (pl["bg"] == rec["planes"]["bg"]).all(), ("bg", t)
                n_c, comps = eng.components(0, t - t0)
                assert sorted(a / 2.0 for a, _ in comps) == rec["areas"], t
                assert sorted(b for _, b in comps) == rec["boxes"], t
                st = stats[0, t - t0]
                assert (bool(st["movement"]), int(st["movement_counter"]), bool(st["wrote"])) == \
                    (rec["movement"], rec["counter"], rec["wrote"])
        return eng.info


def test_cfg4_4k_wide_blur_k193():
    """BASELINE configs[3]: 3840x2160, --box-size 3840 --blur-scale 20 -> k=193 (halo 96)."""
    kw = dict(fps=30, box_size=3840, blur_scale=20, threshold=12, avg=0.1, min_time=0.03, cache_time=0.1)
    info = _oracle_check(3840, 2160, 2, 2, kw, seed=4000)
    assert info["gaussian"] == 193


def test_cfg2_1080p_full_k97_masks():
    from find_motion_b200 import synth
    kw = dict(fps=30, box_size=1920, blur_scale=20, threshold=12, avg=0.1, min_time=0.03, cache_time=0.1,
              mask_areas=synth.CFG2_MASKS)
    info = _oracle_check(1920, 1080, 3, 2, kw, seed=2001)
    assert info["gaussian"] == 97


def test_cfg2_1080p_default_mode_masks():
    from find_motion_b200 import synth
    kw = dict(fps=30, box_size=100, blur_scale=20, threshold=12, avg=0.1, min_time=0.1, cache_time=0.2,
              mask_areas=synth.CFG2_MASKS)
    info = _oracle_check(1920, 1080, 12, 5, kw, seed=2002)
    assert (info["proc_width"], info["proc_height"], info["gaussian"]) == (100, 56, 5)


def test_cfg5_720p_k5():
    kw = dict(fps=30, box_size=1280, blur_scale=256, threshold=12, avg=0.1, min_time=0.1, cache_time=0.2)
    info = _oracle_check(1280, 720, 8, 4, kw, seed=5000)
    assert info["gaussian"] == 5 and info["front_end"] == 0


def test_4k_default_mode():
    kw = dict(fps=30, box_size=100, blur_scale=20, threshold=12, avg=0.1, min_time=0.1, cache_time=0.2)
    _oracle_check(3840, 2160, 3, 3, kw, seed=4001)


@pytest.mark.parametrize("mode", ["full_k5", "default"])
def test_cfg3_64_streams_replicated(mode):
    """64 concurrent 1080p streams in one context: every slot must give the same answer as the
    oracle gives for that clip (slots are replicas of 2 distinct clips), whatever the chunking."""
    import torch
    from find_motion_b200 import synth
    from find_motion_b200.engine import MotionEngine
    from oracle import restated as R
    W, H, n, S = 1920, 1080, 6, 64
    kw = dict(fps=30, threshold=12, avg=0.1, min_time=0.1, cache_time=0.1, mask_areas=synth.CFG2_MASKS)
    kw.update(dict(box_size=1920, blur_scale=384) if mode == "full_k5" else dict(box_size=100, blur_scale=20))
    clips = [synth.make_clip(W, H, n, seed=3000 + i, fps=30, script=[("walker", 1, n), ("blip", 2, 5)]) for i in range(2)]
    want, orcs = [], []
    for c in clips:
        orc = R.StreamOracle(W, H, **kw)
        want.append([orc.process(f) for f in c])
        orcs.append(orc)
    dev = torch.stack([torch.from_numpy(clips[s % 2]) for s in range(S)]).cuda()
    with MotionEngine(W, H, n_streams=S, max_frames=4, **kw) as eng:
        got = np.concatenate([eng.process(dev[:, 0:4]), eng.process(dev[:, 4:6])], axis=1)
        for s in range(S):
            for t in range(n):
                rec = want[s % 2][t]
                st = got[s, t]
                assert int(st["n_contours"]) == len(rec["areas"]), (s, t)
                assert (bool(st["movement"]), int(st["movement_counter"]), int(st["movement_decay"]),
                        bool(st["wrote"]), int(st["n_flush"])) == \
                    (rec["movement"], rec["counter"], rec["decay"], rec["wrote"], rec["n_flush"]), (s, t)
        # the float64 background of replicas is bit-identical, and equals the oracle's after the 6 frames
        bg0 = eng.planes(0, 1, gray=False, blur=False, thresh=False)["bg"]
        bg62 = eng.planes(62, 1, gray=False, blur=False, thresh=False)["bg"]
        bg1 = eng.planes(1, 1, gray=False, blur=False, thresh=False)["bg"]
        assert (bg0 == bg62).all()
        assert (bg0 == orcs[0].bg).all() and (bg1 == orcs[1].bg).all()
        # reset + replay with a different chunking is idempotent
        eng.reset()
        again = np.concatenate([eng.process(dev[:, 0:1]), eng.process(dev[:, 1:4]), eng.process(dev[:, 4:6])], axis=1)
        assert (again == got).all()

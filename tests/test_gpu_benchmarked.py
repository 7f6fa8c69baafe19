"""GPU parity on EXACTLY the configurations bench.py and profiles/run_configs.sh time: the kernels the bench
launches (keep_planes=False instantiations, S x T batches, T = 16 / 32, odd T for the paired bit transpose,
k = 5 / 97 / 385) compared plane by plane with the reference's arithmetic.

Checker: oracle.cv2_chain.Cv2Stream -- the reference's own cv2 call sequence (find_motion.py:487-494, 619-662,
246-276), i.e. the third-party arithmetic the reference runs, at full size in milliseconds per frame; it is
cross-checked against the numpy restatement in tests/test_oracle_vs_cv2.py and on the first frames here.  For
every frame: the dilated threshold plane, contour areas and the decisions; after every call: the float64
background plane.  All for exact equality."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _streams(W, H, n, seeds, fps=30):
    from find_motion_b200 import synth
    script = [("walker", 2, n), ("ring", n // 3, n), ("blip", 1, 4), ("hidden", 0, n)]
    return [synth.make_clip(W, H, n, seed=s, fps=fps, script=script) for s in seeds]


def _check(W, H, S, T, calls, kw, seeds, restated_frames=1, max_components=4096, n_valid=None):
    """S streams (clips cycled over `seeds`), `calls` calls of T frames, keep_planes=False; returns the engine info."""
    cv2 = pytest.importorskip("cv2")
    import torch
    from find_motion_b200.engine import MotionEngine
    from oracle import cv2_chain, restated as R
    n = T * calls
    clips = _streams(W, H, n, seeds, kw.get("fps", 30))
    nd = len(clips)
    refs = [cv2_chain.Cv2Stream(W, H, **kw) for _ in range(nd)]
    restated = R.StreamOracle(W, H, **kw) if restated_frames else None
    dev = torch.stack([torch.from_numpy(clips[s % nd]) for s in range(S)]).cuda()
    with MotionEngine(W, H, n_streams=S, max_frames=T, max_components=max_components, **kw) as eng:
        for c in range(calls):
            stats = eng.process(dev[:, c * T:(c + 1) * T], n_valid=n_valid)
            recs = []
            for d in range(nd):
                rows = []
                for t in range(c * T, (c + 1) * T):
                    rec = refs[d].step(clips[d][t])
                    if d == 0 and restated is not None and t < restated_frames:      # the checker agrees with the restatement
                        r2 = restated.process(clips[d][t], keep_planes=True)
                        assert (r2["planes"]["thresh"] == rec["thresh"]).all() and (r2["planes"]["bg"] == refs[d].ref_frame).all()
                    rows.append(rec)
                recs.append(rows)
            for s in range(S):
                d = s % nd
                for tt in range(T):
                    rec, st = recs[d][tt], stats[s, tt]
                    pl = eng.planes(s, tt, gray=False, blur=False, thresh=True, bg=False)
                    assert (pl["thresh"] == rec["thresh"]).all(), ("thresh", s, c, tt, int((pl["thresh"] != rec["thresh"]).sum()))
                    ncomp, comps = eng.components(s, tt)
                    assert ncomp == len(rec["areas"]) == int(st["n_contours"]), ("contours", s, c, tt)
                    assert sorted(a / 2.0 for a, _ in comps) == rec["areas"], ("areas", s, c, tt)
                    assert (bool(st["movement"]), int(st["movement_counter"]), int(st["movement_decay"]), int(st["cache_len"]),
                            bool(st["wrote"]), int(st["n_flush"])) == \
                        (rec["movement"], rec["counter"], rec["decay"], rec["cache_len"], rec["wrote"], rec["n_flush"]), (s, c, tt)
                bg = eng.planes(s, T - 1, gray=False, blur=False, thresh=False, bg=True)["bg"]
                assert (bg == refs[d].ref_frame).all(), ("bg", s, c, int((bg != refs[d].ref_frame).sum()))
        return eng.info


def _kw(box, blur_scale, masks=True):
    from find_motion_b200 import synth
    kw = dict(fps=30, box_size=box, blur_scale=blur_scale, min_box_scale=50, threshold=12, avg=0.1, min_time=0.5, cache_time=1.0)
    if masks:
        kw["mask_areas"] = synth.CFG2_MASKS
    return kw


def test_bench_headline_1080p_s8_t16_k5():
    """bench.py default: 8 streams x 16 frames, 1920x1080, --blur-scale 384 (k=5), CFG2 masks: k_fused<KEEP=false>."""
    info = _check(1920, 1080, 8, 16, 2, _kw(1920, 384), seeds=[2000, 2001, 2002, 2003])
    assert info["gaussian"] == 5 and info["front_end"] == 0


def test_bench_k97_1080p_s8_t16():
    """other_regimes.full_k97: the reference's own blur scale (--blur-scale 20 at --box-size 1920)."""
    info = _check(1920, 1080, 8, 16, 2, _kw(1920, 20), seeds=[2000, 2001])
    assert info["gaussian"] == 97


def test_bench_default_mode_1080p_s8_t16():
    """other_regimes.default_box100: the reference's CLI default (box 100 -> 100x56, k=5)."""
    info = _check(1920, 1080, 8, 16, 2, _kw(100, 20), seeds=[2000, 2001, 2002, 2003], restated_frames=2)
    assert (info["proc_width"], info["proc_height"], info["gaussian"]) == (100, 56, 5)


def test_cfg4_4k_k385_t16():
    """SURVEY cfg4 extra point: 3840x2160, --blur-scale 10 -> k=385."""
    info = _check(3840, 2160, 1, 16, 1, _kw(3840, 10, masks=False), seeds=[4000], restated_frames=0)
    assert info["gaussian"] == 385


def test_cfg4_4k_k193_t16():
    info = _check(3840, 2160, 1, 16, 1, _kw(3840, 20, masks=False), seeds=[4001], restated_frames=0)
    assert info["gaussian"] == 193


@pytest.mark.parametrize("T", [16, 32])
def test_cfg5_720p_k5_time_blocks(T):
    """SURVEY cfg5 sweep points with the longest time blocks."""
    info = _check(1280, 720, 4, T, 2, _kw(1280, 256, masks=False), seeds=[5000, 5001])
    assert info["gaussian"] == 5 and info["front_end"] == 0


@pytest.mark.parametrize("box,blur_scale", [(1920, 384), (1920, 20)])
def test_odd_time_block_t15(box, blur_scale):
    """T = 15: the last frame of a call has no partner in the paired bit transpose of k_fused."""
    _check(1920, 1080, 2, 15, 2, _kw(box, blur_scale), seeds=[2010, 2011])


def test_single_stream_1080p_t32():
    """profiles single-stream regime: one 1080p stream, T = 32 (the contour stage has 32 frames per call)."""
    _check(1920, 1080, 1, 32, 2, _kw(1920, 384), seeds=[2020])

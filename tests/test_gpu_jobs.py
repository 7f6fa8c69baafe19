"""GPU: the batched job driver (run_pool / run_map / run_stream replacement, find_motion.py:1054-1210) end to end on
the real library: several files of different lengths and frame sizes through one context per geometry, ragged
batches, slot reuse, the pipelined host entry points, and the C-ABI additions they rest on."""
import functools
import io
from collections import deque

import numpy as np
import pytest

from tests.test_jobs_host_logic import KW, MemoryCapture, _clips, expected_writes, make_recorder

pytestmark = pytest.mark.gpu


def test_run_pool_six_files_two_geometries(monkeypatch):
    from find_motion_b200 import jobs
    from find_motion_b200.video_motion import run_vid
    spec = {"a": (96, 72, 31, 11), "b": (96, 72, 12, 12), "c": (160, 120, 25, 13), "d": (96, 72, 40, 14),
            "e": (160, 120, 9, 15), "f": (96, 72, 1, 16), "g": (96, 72, 16, 17)}
    clips = _clips(spec)
    Recorder, written = make_recorder()
    monkeypatch.setattr(jobs, "VideoMotion", Recorder)
    kw = dict(KW, box_size=96, blur_scale=19, mask_areas=[((3, 3), (30, 20)), ((50, 5), (90, 10), (60, 60))])
    job = functools.partial(run_vid, **kw)
    log = io.StringIO()
    res = jobs.run_pool(job, 3, [MemoryCapture(clips[n], n) for n in spec], None, log, devices=[0], streams=3, chunk=5)
    assert sorted(repr(r[1]) for r in res) == sorted(spec)
    for wrote, src, err, objs in res:
        name = repr(src)
        want = expected_writes(clips[name], kw)
        assert err == "" and objs == () and wrote == (len(want) > 0), (name, err)
        got = written.get(name, [])
        assert len(got) == len(want), (name, len(got), len(want))
        for f, t in zip(got, want):
            assert (f == clips[name][t]).all(), (name, t)
    assert len(log.getvalue().strip().splitlines()) == len(spec)


def test_run_pool_on_real_files(tmp_path):
    """Four FFV1 files on disk through cv2.VideoCapture / cv2.VideoWriter: the reference's 4-tuples and exactly the
    frames the reference writes."""
    cv2 = pytest.importorskip("cv2")
    from find_motion_b200 import jobs
    from find_motion_b200.video_motion import run_vid
    spec = {"v0": (256, 192, 40, 51), "v1": (256, 192, 17, 52), "v2": (256, 192, 64, 53), "v3": (256, 192, 5, 54)}
    clips = _clips(spec)
    paths = {}
    for name, clip in clips.items():
        p = str(tmp_path / (name + ".avi"))
        wr = cv2.VideoWriter(p, cv2.VideoWriter_fourcc(*"FFV1"), 6, (256, 192))
        if not wr.isOpened():
            pytest.skip("FFV1 writer not available in this cv2 build")
        for f in clip:
            wr.write(f)
        wr.release()
        paths[p] = name
    cap = cv2.VideoCapture(next(iter(paths)))
    ok, first = cap.read()
    cap.release()
    if not ok or not (first == clips["v0"][0]).all():
        pytest.skip("FFV1 round trip is not lossless here")
    outdir = tmp_path / "out"
    outdir.mkdir()
    kw = dict(KW, box_size=256, blur_scale=51, outdir=str(outdir), codec="FFV1")
    res = jobs.run_pool(functools.partial(run_vid, **kw), 4, list(paths), devices=[0], streams=2, chunk=8)
    assert sorted(r[1] for r in res) == sorted(paths)
    tun = {k: v for k, v in kw.items() if k not in ("outdir", "codec")}
    for wrote, path, err, objs in res:
        name = paths[path]
        want = expected_writes(clips[name], tun)
        assert err == "" and objs == () and wrote == (len(want) > 0)
        outs = sorted(outdir.glob(name + ".avi_*_motion.avi"))
        assert len(outs) == (1 if want else 0)
        if want:
            cap = cv2.VideoCapture(str(outs[0]))
            got = []
            while True:
                ok, f = cap.read()
                if not ok:
                    break
                got.append(f)
            cap.release()
            assert len(got) == len(want)
            for f, t in zip(got, want):
                assert (f == clips[name][t]).all(), (name, t)


def test_run_stream_live_mode_bounded_batches():
    """N4: cameras as sources, chunk 2, deadline per batch; decisions equal the per-stream oracle."""
    from find_motion_b200 import jobs
    from find_motion_b200.video_motion import run_vid
    spec = {"cam0": (96, 72, 21, 61), "cam1": (96, 72, 14, 62), "cam2": (96, 72, 9, 63)}
    clips = _clips(spec)
    Recorder, written = make_recorder()
    import find_motion_b200.jobs as J
    orig = J.VideoMotion
    J.VideoMotion = Recorder
    try:
        res = jobs.run_stream(functools.partial(run_vid, **KW), 3, [MemoryCapture(clips[n], n) for n in spec],
                              io.StringIO(), devices=[0], chunk=2, max_latency=0.2)
    finally:
        J.VideoMotion = orig
    assert sorted(repr(r[1]) for r in res) == sorted(spec)
    for wrote, src, err, objs in res:
        name = repr(src)
        want = expected_writes(clips[name], KW)
        assert err == "" and wrote == (len(want) > 0)
        assert len(written.get(name, [])) == len(want)


def test_ragged_batches_equal_separate_streams():
    """fm_process_ragged: streams advancing by different frame counts per call == each stream alone."""
    import torch
    from find_motion_b200 import synth
    from find_motion_b200.engine import MotionEngine
    from oracle import restated as R
    rng = np.random.default_rng(5)
    for W, H, kw in ((160, 120, dict(fps=6, box_size=160, blur_scale=32, threshold=10, avg=0.2, min_time=0.5, cache_time=1.0)),
                     (160, 120, dict(fps=6, box_size=160, blur_scale=9, threshold=10, avg=0.2, min_time=0.5, cache_time=1.0)),
                     (192, 108, dict(fps=6, box_size=100, blur_scale=20, threshold=8, avg=0.1, min_time=0.3, cache_time=0.5))):
        S, T, n = 4, 6, 30
        clips = np.stack([synth.make_clip(W, H, n, seed=80 + s, fps=6) for s in range(S)])
        orcs = [R.StreamOracle(W, H, **kw) for _ in range(S)]
        pos = [0] * S
        with MotionEngine(W, H, n_streams=S, max_frames=T, **kw) as eng:
            for call in range(9):
                nv = [int(min(rng.integers(0, T + 1), n - pos[s])) for s in range(S)]
                batch = np.zeros((S, T, H, W, 3), np.uint8)
                for s in range(S):
                    batch[s, :nv[s]] = clips[s, pos[s]:pos[s] + nv[s]]
                stats = eng.process(torch.from_numpy(batch).cuda(), n_valid=nv)
                for s in range(S):
                    for t in range(nv[s]):
                        rec = orcs[s].process(clips[s, pos[s] + t], keep_planes=True)
                        st = stats[s, t]
                        assert (int(st["n_contours"]), bool(st["movement"]), int(st["movement_counter"]), int(st["movement_decay"]),
                                int(st["cache_len"]), bool(st["wrote"]), int(st["n_flush"])) == \
                            (len(rec["areas"]), rec["movement"], rec["counter"], rec["decay"], rec["cache_len"], rec["wrote"],
                             rec["n_flush"]), (W, call, s, t)
                        assert (eng.planes(s, t, gray=False, blur=False, bg=False)["thresh"] == rec["planes"]["thresh"]).all()
                    for t in range(nv[s], T):
                        assert tuple(stats[s, t]) == (0,) * 8
                    pos[s] += nv[s]
                    if nv[s]:
                        assert (eng.planes(s, nv[s] - 1, gray=False, blur=False, thresh=False)["bg"] == orcs[s].bg).all(), (W, call, s)


def test_pipelined_host_entry_points_equal_blocking_ones():
    """fm_submit_host / fm_wait on two slots with ragged batches and a mid-run fm_submit_reset == fm_process_host."""
    from find_motion_b200 import synth
    from find_motion_b200.engine import MotionEngine, PinnedBatch
    W, H, S, T, n = 192, 108, 3, 4, 24
    kw = dict(fps=6, box_size=192, blur_scale=38, threshold=10, avg=0.1, min_time=0.3, cache_time=0.6)
    clips = np.stack([synth.make_clip(W, H, n, seed=90 + s, fps=6) for s in range(S)])
    with MotionEngine(W, H, n_streams=S, max_frames=T, **kw) as a, MotionEngine(W, H, n_streams=S, max_frames=T, **kw) as b:
        bufs = [PinnedBatch((S, T, H, W, 3)) for _ in range(3)]
        want, got, pend = [], [], None
        for i, t0 in enumerate(range(0, n, T)):
            if t0 == 12:
                a.reset(1)
                b.submit_reset(1)
            want.append(a.process_host(clips[:, t0:t0 + T]))
            bufs[i % 3].array[:] = clips[:, t0:t0 + T]
            b.submit_host(i & 1, bufs[i % 3].array)
            if pend is not None:
                got.append(b.wait_host(pend))
            pend = i & 1
        got.append(b.wait_host(pend))
        assert all((x == y).all() for x, y in zip(want, got)) and len(want) == len(got) == 6
        assert bufs[0].numa_node >= -1
        for buf in bufs:
            buf.free()


def test_component_records_are_clamped_not_padded():
    """ADVICE: more contours than max_components -> the true count is reported, only real records are returned."""
    import torch
    from find_motion_b200.engine import MotionEngine
    from find_motion_b200 import _lib
    W, H = 256, 128
    frames = np.zeros((1, 2, H, W, 3), np.uint8)
    frames[0, 1, 8::16, 8::16] = 255                      # 8 x 16 = 128 isolated dots in the second frame
    kw = dict(fps=6, box_size=W, blur_scale=W, threshold=5, avg=0.1)
    with MotionEngine(W, H, n_streams=1, max_frames=2, max_components=16, **kw) as eng:
        st = eng.process(torch.from_numpy(frames).cuda())
        n, comps = eng.components(0, 1)
        assert n == int(st[0, 1]["n_contours"]) == 128 and len(comps) == 16
        assert all(a > 0 and w > 0 and h > 0 for a, (x, y, w, h) in comps)
        with pytest.raises(_lib.FmError):
            eng.motion_boxes(0, 1)
    with MotionEngine(W, H, n_streams=1, max_frames=2, max_components=512, **kw) as eng:
        eng.process(torch.from_numpy(frames).cuda())
        assert len(eng.motion_boxes(0, 1)) == 128
        eng.check()


def test_blur_shapes_that_used_to_take_the_naive_kernels():
    """Planes with w % 4 != 0 and k = 1 outside the fused stencil now go through the tensor-core blur."""
    from tests.test_gpu_wide import _run
    for W, H, bs, fe in ((111, 67, 11, 1), (333, 217, 30, 1), (150, 40, 150, 1)):
        kw = dict(fps=6, box_size=W, blur_scale=bs, threshold=5, avg=0.2, min_time=0.3, cache_time=0.6,
                  mask_areas=[((2, 1), (W // 3, H // 2))])
        _run(W, H, 7, 3, kw, seed=700 + W, expect_front_end=fe)
    # resize to an odd width (box 111 from 640x480): k = 5 on a 111-pixel plane
    kw = dict(fps=6, box_size=111, blur_scale=20, threshold=6, avg=0.15, min_time=0.3, cache_time=0.5)
    _run(640, 480, 8, 4, kw, seed=720, expect_front_end=2)


def test_detector_input_plane_matches_cv2():
    """N3 (find_motion.py:703-706): imutils.resize(frame.raw, width=300) on the GPU == cv2.resize(INTER_AREA)."""
    cv2 = pytest.importorskip("cv2")
    from find_motion_b200 import synth
    from find_motion_b200.engine import resize_area
    for W, H, width in ((1920, 1080, 300), (640, 480, 300), (1280, 720, 300), (600, 338, 300), (333, 217, 100), (300, 200, 300)):
        frame = synth.make_clip(W, H, 1, seed=W + width)[0]
        want = cv2.resize(frame, (width, int(H * (width / float(W)))), interpolation=cv2.INTER_AREA)
        got = resize_area(frame, width)
        assert got.shape == want.shape and (got == want).all(), (W, H, width, int((got != want).sum()))

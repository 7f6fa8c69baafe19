"""CPU: host logic of the batched job driver (find_motion_b200/jobs.py, the replacement of run_pool / run_map /
run_stream, find_motion.py:1054-1210) with the device replaced by a stand-in that answers every batch from the
oracle: slot assignment and reuse, ragged batches, reset ordering, replay of decide_output on the raw frames,
result tuples, progress log, error reporting, mixed geometries.  No GPU, no libfmgpu compute calls."""
import functools
import io
from collections import deque

import numpy as np
import pytest

from find_motion_b200 import synth
from find_motion_b200.engine import STATS_DTYPE
from oracle import restated as R


class MemoryCapture:
    """cv2.VideoCapture stand-in (get/read/isOpened/release)."""

    def __init__(self, frames, name):
        self.frames, self.i, self.name = frames, 0, name

    def get(self, prop):
        import cv2
        return {cv2.CAP_PROP_FRAME_COUNT: float(len(self.frames)), cv2.CAP_PROP_FRAME_WIDTH: float(self.frames.shape[2]),
                cv2.CAP_PROP_FRAME_HEIGHT: float(self.frames.shape[1])}.get(prop, 0.0)

    def isOpened(self):
        return True

    def read(self):
        if self.i >= len(self.frames):
            return False, None
        self.i += 1
        return True, self.frames[self.i - 1]

    def release(self):
        pass

    def __repr__(self):
        return self.name


class OracleEngine:
    """MotionEngine stand-in: the slots are StreamOracle instances; batches are answered at wait time."""
    instances = []

    def __init__(self, W, H, n_streams=1, max_frames=8, device=0, mask_areas=None, **kw):
        self.W, self.H, self.S, self.T, self.kw, self.masks = W, H, n_streams, max_frames, kw, mask_areas
        p = R.derive_params(W, H, kw.get("fps", 30), kw.get("box_size", 100), kw.get("min_box_scale", 50),
                            kw.get("cache_time", 2.0), kw.get("min_time", 0.5), kw.get("blur_scale", 20))
        self.info = dict(scale=p["scale"], max_area=p["max_area"], min_area=p["min_area"], gaussian=p["gaussian"])
        self.slots = [None] * n_streams
        self.queue = {}
        self.batches, self.ragged, self.resets = 0, 0, 0
        self.closed = False
        OracleEngine.instances.append(self)

    def _fresh(self):
        return R.StreamOracle(self.W, self.H, mask_areas=self.masks, **self.kw)

    def submit_reset(self, s):
        self.resets += 1
        self.queue.setdefault("ops", []).append(("reset", s))

    def submit_host(self, slot, host, n_valid=None):
        assert slot not in self.queue, "slot reused before wait"
        nv = list(n_valid) if n_valid is not None else [host.shape[1]] * self.S
        self.queue.setdefault("ops", []).append(("batch", slot, host, nv))
        self.queue[slot] = True

    def wait_host(self, slot):
        ops = self.queue["ops"]
        out = None
        while ops:                           # stream order: resets and batches as submitted
            op = ops.pop(0)
            if op[0] == "reset":
                self.slots[op[1]] = self._fresh()
                continue
            _, bslot, host, nv = op
            assert bslot == slot, "batches complete in submission order"
            T = host.shape[1]
            stats = np.zeros((self.S, T), STATS_DTYPE)
            self.batches += 1
            self.ragged += int(len(set(nv)) > 1)
            for s in range(self.S):
                for t in range(nv[s]):
                    rec = self.slots[s].process(host[s, t].copy())
                    stats[s, t] = (len(rec["areas"]), len(rec["areas"]), rec["movement"], rec["counter"], rec["decay"],
                                   rec["cache_len"], rec["wrote"], rec["n_flush"])
            out = stats
            break
        del self.queue[slot]
        return out

    def close(self):
        self.closed = True


class HostBuffer:
    def __init__(self, shape, device=0):
        self.array = np.zeros(shape, np.uint8)
        self.numa_node = -1

    def free(self):
        self.array = None


def make_recorder():
    from find_motion_b200.video_motion import VideoMotion
    written = {}

    class Recorder(VideoMotion):
        def _make_outfile(self):
            self.outfiles += 1
            name = repr(self.filename)

            class Wr:
                def write(_, frame):
                    written.setdefault(name, []).append(frame.copy())

                def release(_):
                    pass
            self.outfile = Wr()
    return Recorder, written


def expected_writes(clip, kw):
    orc = R.StreamOracle(clip.shape[2], clip.shape[1], **kw)
    out, cache = [], deque(maxlen=orc.p["cache_frames"])
    for t, f in enumerate(clip):
        rec = orc.process(f)
        if rec["wrote"]:
            if rec["n_flush"]:
                out += list(cache)
                cache.clear()
            out.append(t)
        else:
            cache.append(t)
    return out


KW = dict(fps=6, box_size=96, blur_scale=19, threshold=10, avg=0.2, min_time=0.4, cache_time=0.7, min_box_scale=50)


def _clips(spec):
    return {name: synth.make_clip(W, H, n, seed=seed, fps=6) for name, (W, H, n, seed) in spec.items()}


@pytest.mark.parametrize("streams,chunk", [(2, 5), (3, 4), (8, 16)])
def test_scheduler_matches_per_stream_oracle(streams, chunk):
    from find_motion_b200 import jobs
    from find_motion_b200.video_motion import run_vid
    spec = {"a": (96, 72, 31, 11), "b": (96, 72, 12, 12), "c": (96, 72, 5, 13), "d": (96, 72, 40, 14),
            "e": (96, 72, 20, 15), "f": (96, 72, 1, 16), "g": (96, 72, 16, 17)}
    clips = _clips(spec)
    Recorder, written = make_recorder()
    OracleEngine.instances.clear()
    job = functools.partial(run_vid, **KW)
    sched = jobs.BatchScheduler(job, devices=[0], streams=streams, chunk=chunk, stream_cls=Recorder,
                                engine_factory=OracleEngine, buffer_factory=HostBuffer)
    sources = [MemoryCapture(clips[n], n) for n in spec]
    seen = []
    res = sched.run(sources, seen.append)
    assert sorted(repr(r[1]) for r in res) == sorted(spec) and res == seen
    for wrote, src, err, objs in res:
        name = repr(src)
        want = expected_writes(clips[name], KW)
        assert err == "" and objs == ()
        assert wrote == (len(want) > 0), name
        got = written.get(name, [])
        assert len(got) == len(want), (name, len(got), len(want))
        for f, t in zip(got, want):
            assert (f == clips[name][t]).all(), (name, t)
    eng = OracleEngine.instances[0]
    assert len(OracleEngine.instances) == 1 and eng.closed
    assert eng.resets == len(spec)                       # every stream starts from a reset slot
    if streams < len(spec):
        assert eng.ragged > 0                            # streams of different lengths made ragged batches


def test_run_pool_mixed_geometries_errors_and_progress_log(tmp_path, monkeypatch):
    """run_pool's contract (find_motion.py:1054-1122): one tuple per input, errors reported not raised, progress-log
    lines only for the inputs without error; inputs of another frame size get a context of their own."""
    from find_motion_b200 import jobs
    from find_motion_b200.video_motion import run_vid
    spec = {"s1": (96, 72, 9, 21), "big1": (128, 64, 14, 22), "s2": (96, 72, 18, 23), "big2": (128, 64, 3, 24)}
    clips = _clips(spec)
    Recorder, written = make_recorder()
    OracleEngine.instances.clear()
    monkeypatch.setattr(jobs, "MotionEngine", OracleEngine)
    monkeypatch.setattr(jobs, "PinnedBatch", HostBuffer)
    monkeypatch.setattr(jobs, "VideoMotion", Recorder)
    kw = dict(KW, box_size=64, blur_scale=13)
    job = functools.partial(run_vid, **kw)
    log = io.StringIO()

    class Bar:
        seen = []

        def update(self, n):
            self.seen.append(n)
    missing = str(tmp_path / "missing.avi")
    sources = [MemoryCapture(clips[n], n) for n in spec] + [missing]
    res = jobs.run_pool(job, 4, sources, Bar(), log, devices=[0], streams=2, chunk=4)
    assert len(res) == 5 and Bar.seen == [1, 2, 3, 4, 5]
    by_name = {repr(r[1]) if not isinstance(r[1], str) else r[1]: r for r in res}
    assert by_name[missing][0] is None and by_name[missing][2] == 'Video did not load successfully'
    for name in spec:
        wrote, _, err, objs = by_name[name]
        want = expected_writes(clips[name], kw)
        assert err == "" and wrote == (len(want) > 0)
        assert [int((f == clips[name][t]).all()) for f, t in zip(written.get(name, []), want)] == [1] * len(want)
    lines = log.getvalue().strip().splitlines()
    assert sorted(l.split(" // ")[0] for l in lines) == sorted(spec) and all(l.endswith(" // ()") for l in lines)
    assert sorted((e.W, e.H) for e in OracleEngine.instances) == [(96, 72), (128, 64)]      # one context per geometry


def test_run_map_keeps_input_order_and_run_stream_reports(monkeypatch):
    from find_motion_b200 import jobs
    from find_motion_b200.video_motion import run_vid
    spec = {"m1": (96, 72, 7, 31), "m2": (96, 72, 9, 32), "m3": (96, 72, 4, 33)}
    clips = _clips(spec)
    Recorder, _ = make_recorder()
    monkeypatch.setattr(jobs, "MotionEngine", OracleEngine)
    monkeypatch.setattr(jobs, "PinnedBatch", HostBuffer)
    monkeypatch.setattr(jobs, "VideoMotion", Recorder)
    job = functools.partial(run_vid, **KW)
    res = jobs.run_map(job, [MemoryCapture(clips[n], n) for n in spec], chunk=4)
    assert [repr(r[1]) for r in res] == list(spec)
    with pytest.raises(ValueError):
        jobs.run_map(job, [])
    with pytest.raises(ValueError):
        jobs.run_pool(job, 2, [])
    # live mode: small batches closed by a deadline; the "cameras" here deliver frames without pacing
    log = io.StringIO()
    OracleEngine.instances.clear()
    res = jobs.run_stream(job, 3, [MemoryCapture(clips[n], n) for n in spec], log, devices=[0], chunk=2, max_latency=0.05)
    assert sorted(repr(r[1]) for r in res) == sorted(spec) and all(r[2] == "" for r in res)
    assert log.getvalue().count("Finished streaming from camera") == 3
    assert OracleEngine.instances[0].S == 3 and OracleEngine.instances[0].T == 2


def test_pause_event_holds_the_drivers(monkeypatch):
    """find_motion.py:124-170, 858-860: a cleared `unpaused` Event stops the stream loop until it is set again."""
    import threading
    import time
    from find_motion_b200 import jobs
    from find_motion_b200.video_motion import run_vid
    clips = _clips({"p": (96, 72, 6, 41)})
    Recorder, _ = make_recorder()
    monkeypatch.setattr(jobs, "MotionEngine", OracleEngine)
    monkeypatch.setattr(jobs, "PinnedBatch", HostBuffer)
    monkeypatch.setattr(jobs, "VideoMotion", Recorder)
    job = functools.partial(run_vid, **KW)
    jobs.unpaused.clear()
    box = {}
    th = threading.Thread(target=lambda: box.update(res=jobs.run_map(job, [MemoryCapture(clips["p"], "p")], chunk=3)))
    th.start()
    time.sleep(0.3)
    assert th.is_alive() and "res" not in box
    jobs.unpaused.set()
    th.join(timeout=60)
    assert len(box["res"]) == 1 and box["res"][0][2] == ""

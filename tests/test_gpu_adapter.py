"""GPU: the VideoMotion / run_vid drop-in writes exactly the frames the reference writes."""
from collections import deque

import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu


class MemoryCapture:
    """cv2.VideoCapture stand-in (get/read/isOpened/release)."""

    def __init__(self, frames):
        self.frames, self.i = frames, 0

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


def expected_written(trace, cache_frames):
    """Replay decide_output (find_motion.py:549-589) on frame indices from the golden decisions."""
    out, cache = [], deque(maxlen=cache_frames)
    for t, e in enumerate(trace):
        if e["wrote"]:
            if e["n_flush"]:
                out += list(cache)
                cache.clear()
            out.append(t)
        else:
            cache.append(t)
    return out


@pytest.mark.parametrize("name,chunk", [("cfg1_640x480_default", 16), ("full_256x192_k5", 5),
                                        ("full_320x240_k17_masks", 32)])
def test_adapter_writes_reference_frames(name, chunk):
    pytest.importorskip("cv2")
    from find_motion_b200.video_motion import VideoMotion
    fx = helpers.load_golden(name)
    clip = helpers.golden_clip(fx)
    # tag every frame so that written frames can be identified
    written = []

    class Recorder(VideoMotion):
        def _make_outfile(self):
            self.outfiles += 1

            class W:
                def write(_, frame):
                    written.append(frame.copy())      # the frame lies in a pinned batch that is reused

                def release(_):
                    pass
            self.outfile = W()

    vm = Recorder(MemoryCapture(clip), chunk=chunk, **fx["kwargs"])
    assert vm.loaded
    assert vm.cache_frames == fx["params"]["cache_frames"] and vm.max_area == fx["params"]["max_area"]
    wrote, err, seen = vm.find_motion()
    assert err == "" and seen == ()
    assert wrote == fx["result"][0]
    want = expected_written(fx["trace"], fx["params"]["cache_frames"])
    assert len(written) == len(want) == fx["writes"]
    for frame, t in zip(written, want):
        assert frame is clip[t] or (frame == clip[t]).all()


def test_run_vid_reports_errors_like_the_reference():
    from find_motion_b200.video_motion import run_vid
    res = run_vid(None)
    assert res[0] is None and "Filename required" in res[2]
    res = run_vid("/nonexistent/file.avi", box_size=100)
    assert res[0] is None and res[2] != ""


def test_run_vid_on_a_real_video_file(tmp_path):
    """End to end through cv2.VideoCapture / cv2.VideoWriter: a lossless FFV1 clip in, *_motion.avi out
    with exactly the frames the reference writes (golden trace of the same clip)."""
    cv2 = pytest.importorskip("cv2")
    from find_motion_b200.video_motion import run_vid
    fx = helpers.load_golden("full_256x192_k5")
    clip = helpers.golden_clip(fx)
    W, H = fx["clip"]["W"], fx["clip"]["H"]
    src = str(tmp_path / "clip.avi")
    wr = cv2.VideoWriter(src, cv2.VideoWriter_fourcc(*"FFV1"), fx["kwargs"]["fps"], (W, H))
    if not wr.isOpened():
        pytest.skip("FFV1 writer not available in this cv2 build")
    for f in clip:
        wr.write(f)
    wr.release()
    cap = cv2.VideoCapture(src)
    ok, first = cap.read()
    cap.release()
    if not ok or not (first == clip[0]).all():
        pytest.skip("FFV1 round trip is not lossless here")
    outdir = tmp_path / "out"
    outdir.mkdir()
    kw = dict(fx["kwargs"], outdir=str(outdir), codec="FFV1", chunk=8)
    wrote, name, err, seen = run_vid(src, **kw)
    assert err == "" and wrote is True and name == src and seen == ()
    outs = sorted(outdir.glob("*_motion.avi"))
    assert len(outs) == 1
    cap = cv2.VideoCapture(str(outs[0]))
    got = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        got.append(f)
    cap.release()
    want = expected_written(fx["trace"], fx["params"]["cache_frames"])
    assert len(got) == len(want) == fx["writes"]
    for f, t in zip(got, want):
        assert (f == clip[t]).all()


def test_motion_boxes_match_the_reference_overlay():
    """the boxes --show would draw = boundingRect scaled by 1/scale (find_motion.py:787-813)."""
    import torch
    from find_motion_b200.engine import MotionEngine
    fx = helpers.load_golden("cfg2_1080p_D_masks")
    clip = helpers.golden_clip(fx)[:24]
    c = fx["clip"]
    with MotionEngine(c["W"], c["H"], n_streams=1, max_frames=8, **fx["kwargs"]) as eng:
        dev = torch.from_numpy(clip).cuda()
        inv = 1 / fx["params"]["scale"]
        for t0 in range(0, 24, 8):
            eng.process(dev[None, t0:t0 + 8])
            for t in range(t0, t0 + 8):
                want = sorted(((int(x * inv), int(y * inv)), (int((x + w) * inv), int((y + h) * inv)))
                              for (x, y, w, h) in fx["trace"][t]["boxes"])
                assert eng.motion_boxes(0, t - t0) == want

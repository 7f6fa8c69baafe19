"""TEST INFRASTRUCTURE ONLY -- loads the *real* reference (find_motion.py) as an oracle.

This module is the "reference-exec" half of the oracle.
It imports /root/reference/find_motion/find_motion.py unmodified, with stub modules for
its non-arithmetic dependencies that are absent from this image (pynput, progressbar,
mem_top, orderedset, cvlib, imutils), and drives the reference's own
``VideoMotion.find_motion()`` loop (find_motion.py:852-904) from in-memory frames while
recording every plane and decision.

It only works where /root/reference exists (the build container).  It is used by
  * tests/golden/make_golden.py   -- to generate the committed golden fixtures, and
  * tests/test_oracle_vs_cv2.py::test_against_live_reference_loop -- skipped when the reference is absent.
Nothing in the product package imports it.  It cannot travel to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import sys
import types
from collections import deque

import numpy as np

REFERENCE_DIR = os.environ.get("FM_REFERENCE_DIR", "/root/reference/find_motion")


def reference_available() -> bool:
    if not os.path.isfile(os.path.join(REFERENCE_DIR, "find_motion.py")):
        return False
    try:
        import cv2  # noqa: F401
    except Exception:
        return False
    return True


_fm = None


def load_reference():
    """Import the reference module with stubs (Appendix B recipe) and return it."""
    global _fm
    if _fm is not None:
        return _fm
    import cv2

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class Key:
        esc = "esc"
        pause = "pause"
        shift = "shift"
        alt_l = "alt_l"

    class Listener:
        def __init__(self, *a, **k):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def stop(self):
            pass

    kb = stub("pynput.keyboard", Key=Key, Listener=Listener)
    stub("pynput", keyboard=kb)
    stub("progressbar", ProgressBar=type("ProgressBar", (), {"__init__": lambda s, *a, **k: None}))
    stub("mem_top", mem_top=lambda: "")
    stub("orderedset", OrderedSet=list)
    stub("cvlib", detect_common_objects=lambda *a, **k: ([], [], []))

    def imutils_resize(image, width=None, height=None, inter=cv2.INTER_AREA):
        (h, w) = image.shape[:2]
        if width is None and height is None:
            return image
        if width is None:
            r = height / float(h)
            dim = (int(w * r), height)
        else:
            r = width / float(w)
            dim = (width, int(h * r))
        return cv2.resize(image, dim, interpolation=inter)

    stub("imutils", resize=imutils_resize)
    cv2.waitKey = lambda delay=0: -1  # headless cv2 raises in waitKey (find_motion.py:821/894)
    # the reference does a bare "from DummyProgressBar import ..." (find_motion.py:55); an
    # already-imported top-level package called find_motion would shadow it, so load by path.
    import importlib.util

    sys.path.insert(0, REFERENCE_DIR)
    try:
        spec = importlib.util.spec_from_file_location(
            "_reference_find_motion", os.path.join(REFERENCE_DIR, "find_motion.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["_reference_find_motion"] = mod
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(REFERENCE_DIR)
    _fm = mod
    return mod


class MemoryCapture:
    """Stands in for cv2.VideoCapture (find_motion.py:413, 501): frames come from a list."""

    def __init__(self, frames, width, height):
        self.frames = frames
        self.i = 0
        self.width = width
        self.height = height

    def get(self, prop):
        import cv2

        if prop == cv2.CAP_PROP_FRAME_COUNT:
            return float(len(self.frames))
        if prop == cv2.CAP_PROP_FRAME_WIDTH:
            return float(self.width)
        if prop == cv2.CAP_PROP_FRAME_HEIGHT:
            return float(self.height)
        return 0.0

    def isOpened(self):
        return True

    def read(self):
        if self.i >= len(self.frames):
            return False, None
        f = self.frames[self.i]
        self.i += 1
        return True, f

    def release(self):
        pass


def sha(a: np.ndarray) -> str:
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_reference(frames, keep_planes=False, **kwargs):
    """Run the reference's VideoMotion loop over in-memory BGR frames.

    Returns dict with 'params' (derived), 'trace' (per-frame dicts) and 'result'
    (the tuple find_motion() returned).  Each trace entry holds sha1 hashes of
    gray / blur(masked) / thresh(dilated) / bg(float64), sorted contour areas and
    bounding boxes, and the counters after decide_output.
    """
    import cv2

    fm = load_reference()
    frames = list(frames)
    H, W = frames[0].shape[:2]
    trace = []

    class Recorder(fm.VideoMotion):
        def _load_video(self):  # find_motion.py:409-424 with an in-memory capture
            self.cap = MemoryCapture(frames, W, H)
            self.ref_frame = None
            self.frame_cache = deque(maxlen=self.cache_frames)
            self._get_video_info()
            self.scale = self.box_size / self.frame_width
            self.max_area = int((self.frame_width * self.frame_height) / 2 * self.scale)
            self._writes = 0
            return True

        def output_raw_frame(self, frame=None):  # find_motion.py:533-546 without the file
            self.wrote_frames = True
            self._writes += 1
            self._flushed += 1

        def output_frame(self, frame=None):  # find_motion.py:509-530 without the file
            self.wrote_frames = True
            self._writes += 1
            self._wrote_current = True

        def find_objects(self, *a, **k):  # out of scope
            return set()

        def decide_output(self):
            cf = self.current_frame
            entry = {
                "gray": sha(cf.gray),
                "blur": sha(cf.blur),
                "thresh": sha(cf.thresh),
                "bg": sha(self.ref_frame),
                "areas": sorted(float(cv2.contourArea(c)) for c in cf.contours),
                "boxes": sorted(tuple(int(v) for v in cv2.boundingRect(c)) for c in cf.contours),
            }
            if keep_planes:
                entry["planes"] = {
                    "gray": cf.gray.copy(), "blur": cf.blur.copy(),
                    "thresh": cf.thresh.copy(), "bg": self.ref_frame.copy(),
                }
            self._flushed = 0
            self._wrote_current = False
            super().decide_output()
            entry.update(
                movement=bool(self.movement), counter=int(self.movement_counter),
                decay=int(self.movement_decay), cache_len=len(self.frame_cache),
                wrote=bool(self._wrote_current), n_flush=int(self._flushed),
            )
            trace.append(entry)

    vm = Recorder(filename="memory", **kwargs)
    params = {
        "scale": vm.scale, "gaussian": vm.gaussian[0], "min_area": vm.min_area,
        "max_area": vm.max_area, "cache_frames": vm.cache_frames,
        "min_movement_frames": vm.min_movement_frames,
    }
    result = vm.find_motion()
    return {"params": params, "trace": trace, "result": result, "writes": vm._writes}

"""TEST / BASELINE INFRASTRUCTURE ONLY -- the reference's per-frame loop re-typed as the same
sequence of cv2 calls it makes, for timing the CPU path on the GPU box (where /root/reference
does not exist) and for cross-checking the numpy restatement.  Nothing in find_motion_b200/
imports this.  Each step cites the reference line it mirrors (find_motion/find_motion.py).

This is a "port" in bench.py's vocabulary: same third-party arithmetic (opencv-python), same
order of calls, same per-stream process fan-out (multiprocessing.Pool, one stream per task,
find_motion.py:1071-1075), with decode, encode, display and object detection left out exactly
as in the GPU measurement.
"""
from __future__ import annotations

import math
import time
from collections import deque

import numpy as np


def available() -> bool:
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False


class Cv2Stream:
    """VideoMotion's state and per-frame work (find_motion.py:299-380, 852-904) without I/O."""

    def __init__(self, W, H, fps=30, box_size=100, min_box_scale=50, cache_time=2.0, min_time=0.5,
                 threshold=7, avg=0.1, blur_scale=20, mask_areas=None):
        self.box_size = box_size
        self.cache_frames = int(cache_time * fps)                 # :334
        self.min_movement_frames = int(min_time * fps)            # :335
        self.delta_thresh = threshold
        self.avg = avg
        self.mask_areas = mask_areas if mask_areas is not None else []
        self.min_area = int(math.pow(box_size / min_box_scale, 2))            # :406
        g = int(box_size / blur_scale)                            # :482
        g = g + 1 if g % 2 == 0 else g                            # :483
        self.gaussian = (g, g)
        self.scale = box_size / W                                 # :422
        self.max_area = int((W * H) / 2 * self.scale)             # :423
        self.ref_frame = None
        self.frame_cache = deque(maxlen=self.cache_frames)        # :415
        self.movement = False
        self.movement_decay = 0
        self.movement_counter = 0
        self.writes = 0

    def step(self, raw):
        import cv2
        frame_copy = raw.copy()                                   # VideoFrame.__init__, :236
        (h, w) = raw.shape[:2]                                    # imutils.resize(width=box_size), :492
        r = self.box_size / float(w)
        small = cv2.resize(raw, (self.box_size, int(h * r)), interpolation=cv2.INTER_AREA)
        gray = cv2.cvtColor(small, cv2.COLOR_BGR2GRAY)            # :493
        blur = cv2.GaussianBlur(gray, self.gaussian, 0)           # :494
        for area in self.mask_areas:                              # :626-635
            pts = [(int(a[0] * self.scale), int(a[1] * self.scale)) for a in area]
            if len(pts) == 2:
                cv2.rectangle(blur, *pts, (0, 0, 0), cv2.FILLED)
            else:
                cv2.fillConvexPoly(blur, np.array(pts, np.int32), (0, 0, 0))
        if self.ref_frame is None:                                # :651-652
            self.ref_frame = blur.copy().astype("float")
        delta = cv2.absdiff(blur, cv2.convertScaleAbs(self.ref_frame))                      # :250
        thresh = cv2.threshold(delta, self.delta_thresh, maxval=255, type=cv2.THRESH_BINARY)[1]   # :257
        cv2.accumulateWeighted(blur, self.ref_frame, self.avg)    # :659
        thresh = cv2.dilate(thresh, kernel=None, iterations=2)    # :266
        cnts = cv2.findContours(thresh, mode=cv2.RETR_EXTERNAL, method=cv2.CHAIN_APPROX_SIMPLE)[-2]   # :269-272
        self.movement = False                                     # :671
        self.movement_decay -= 1 if self.movement_decay > 0 else 0    # :672
        areas = []
        for c in cnts:                                            # :676
            area = cv2.contourArea(c)                             # :679
            areas.append(area)
            if self.max_area < area < self.min_area:              # :684
                continue
            self.movement_counter += 1                            # :694
            self.movement = True
        if not self.movement:                                     # :697
            self.movement_counter = 0
        n_flush, wrote = 0, False
        if self.movement_counter >= self.min_movement_frames or self.movement_decay > 0:    # :555
            if self.movement:
                self.movement_decay = self.cache_frames           # :559
                n_flush = len(self.frame_cache)
                self.frame_cache.clear()                          # :570
            wrote = True                                          # :583
            self.writes += 1 + n_flush
        else:
            self.frame_cache.append(frame_copy)                   # :588
        return {"areas": sorted(areas), "movement": self.movement, "counter": self.movement_counter,
                "decay": self.movement_decay, "cache_len": len(self.frame_cache), "wrote": wrote,
                "n_flush": n_flush, "thresh": thresh, "blur": blur, "gray": gray}


# ------------------------------------------------------------------------------------------
# timing fan-out: one stream per worker process, like run_pool (find_motion.py:1054-1122).
# Workers are plain subprocesses (`python -m oracle.cv2_chain <json>`): forking a parent that has
# already initialised cv2's thread pool or CUDA deadlocks, and spawn needs an importable main.
# ------------------------------------------------------------------------------------------

def _worker(spec):
    from find_motion_b200 import synth
    W, H, kw = spec["W"], spec["H"], spec["kw"]
    if kw.get("mask_areas"):
        kw["mask_areas"] = [tuple(tuple(p) for p in a) for a in kw["mask_areas"]]
    steps, warmup = spec.get("steps", 1), spec.get("warmup", 0)
    out = []
    for seed in spec["seeds"]:
        clip = synth.make_clip(W, H, spec["clip_len"], seed, fps=kw.get("fps", 30))
        if spec["use_cv2"]:
            import cv2
            if spec["cv_threads"] > 0:            # 0 = the reference's implicit default (cv2 threads = cores)
                cv2.setNumThreads(spec["cv_threads"])
            step = Cv2Stream(W, H, **kw).step
        else:
            from oracle import restated as R
            step = R.StreamOracle(W, H, **kw).process
        if spec.get("source") == "ffv1" and spec["use_cv2"]:
            # decode included: the clip is written once as a lossless FFV1 file (untimed) and the timed loop pulls its
            # frames through cv2.VideoCapture.read(), re-opening the file when it is exhausted (find_motion.py:497-506)
            import os
            import tempfile
            path = os.path.join(tempfile.mkdtemp(prefix="fm_cpu_"), "clip_%d.avi" % seed)
            wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), kw.get("fps", 30), (W, H))
            for f in clip:
                wr.write(f)
            wr.release()
            state = {"cap": cv2.VideoCapture(path)}

            def next_frame(i):
                ok, f = state["cap"].read()
                if not ok:
                    state["cap"].release()
                    state["cap"] = cv2.VideoCapture(path)
                    ok, f = state["cap"].read()
                return f
        else:
            def next_frame(i):
                return clip[i % spec["clip_len"]]
        moved, i = 0, 0
        for _ in range(warmup * spec["frames"]):
            step(next_frame(i))
            i += 1
        t0 = time.perf_counter()
        for _ in range(steps * spec["frames"]):
            moved += int(step(next_frame(i))["movement"])
            i += 1
        out.append({"seconds": time.perf_counter() - t0, "moved": moved})
    return out


def time_cpu_path(W, H, kw, n_streams, frames_per_stream, processes, clip_len=8, seed0=2000, cv_threads=1,
                  timeout=900, steps=1, warmup=0, source="memory"):
    """Run `n_streams` synthetic streams over `processes` worker processes (streams dealt round-robin);
    every stream does `warmup` untimed + `steps` timed steps of `frames_per_stream` frames.
    Returns dict(fps, seconds, frames, processes, engine)."""
    import json
    import os
    import subprocess
    import sys
    use_cv2 = available()
    processes = max(1, min(processes, n_streams))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    procs = []
    for p in range(processes):
        spec = {"W": W, "H": H, "kw": kw, "seeds": [seed0 + s for s in range(p, n_streams, processes)],
                "clip_len": clip_len, "frames": frames_per_stream, "cv_threads": cv_threads, "use_cv2": use_cv2,
                "steps": steps, "warmup": warmup, "source": source}
        procs.append(subprocess.Popen([sys.executable, "-m", "oracle.cv2_chain", json.dumps(spec)], cwd=root,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    busy = 0.0
    for pr in procs:
        so, se = pr.communicate(timeout=timeout)
        if pr.returncode != 0:
            raise RuntimeError("cpu worker failed: " + se[-500:])
        res = json.loads(so.strip().splitlines()[-1])
        busy = max(busy, sum(r["seconds"] for r in res))     # a worker runs its streams back to back
    frames = n_streams * frames_per_stream * steps
    return {"fps": frames / busy, "seconds": busy, "frames": frames, "processes": processes,
            "engine": "cv2 call chain" if use_cv2 else "numpy oracle", "cv_threads": cv_threads}


if __name__ == "__main__":
    import json
    import sys
    print(json.dumps(_worker(json.loads(sys.argv[1]))))

"""Generate tests/golden/*.json from the REAL reference (find_motion.py run via oracle/ref_loader).

Run in the build container only (needs /root/reference and cv2):
    python tests/golden/make_golden.py
Each fixture holds, per frame, sha1 hashes of the reference's gray / blur(masked) /
thresh(dilated) / float64 background planes, the sorted contour areas and bounding boxes, and
the movement / counter / decay / cache / output decisions, for a seeded synthetic clip that the
tests regenerate with find_motion_b200.synth.make_clip (same numpy version on the GPU box).
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from find_motion_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

README_MASKS = synth.README_MASKS
CFG2_MASKS = synth.CFG2_MASKS

# name -> (clip spec, VideoMotion kwargs).  CLI defaults are threshold 12, cachetime 1.0,
# mintime 0.5, avg 0.1, blur_scale 20, box_size 100, min_box_scale 50 (find_motion.py:1452-1489)
CLI = dict(threshold=12, cache_time=1.0, min_time=0.5, avg=0.1, blur_scale=20, box_size=100,
           min_box_scale=50)
CASES = {
    # BASELINE.json configs[0]: 640x480 15 fps, defaults, no masks (100x75, k=5, 12-element tail)
    "cfg1_640x480_default": (dict(W=640, H=480, n=300, seed=synth.stream_seed(1, 0), fps=15),
                             dict(CLI, fps=15)),
    # configs[1] in default mode (box 100 -> 100x56) with the README + translated masks
    "cfg2_1080p_D_masks": (dict(W=1920, H=1080, n=64, seed=synth.stream_seed(2, 0), fps=10),
                           dict(CLI, fps=10, mask_areas=CFG2_MASKS)),
    # full-res mode, k=17, README masks
    "full_320x240_k17_masks": (dict(W=320, H=240, n=96, seed=synth.stream_seed(2, 1), fps=8),
                               dict(CLI, fps=8, box_size=320, mask_areas=README_MASKS)),
    # full-res mode, k=5 (the HBM-bound regime), other threshold / avg
    "full_256x192_k5": (dict(W=256, H=192, n=96, seed=synth.stream_seed(5, 0), fps=8),
                        dict(CLI, fps=8, box_size=256, blur_scale=64, threshold=7, avg=0.3)),
    # integer-ratio resize paths: 2x2 and 4x4
    "half_640x480_box320": (dict(W=640, H=480, n=48, seed=synth.stream_seed(1, 1), fps=4),
                            dict(CLI, fps=4, box_size=320)),
    "quarter_640x480_box160": (dict(W=640, H=480, n=48, seed=synth.stream_seed(1, 2), fps=4),
                               dict(CLI, fps=4, box_size=160)),
    # odd sizes: width not a multiple of 32, N % 16 != 0 tail, min_time 0 / cache 0 edge cases
    "odd_333x217_box111": (dict(W=333, H=217, n=64, seed=7, fps=6),
                           dict(CLI, fps=6, box_size=111, blur_scale=10, min_time=0.0, cache_time=0.0)),
    "odd_full_203x117_k9": (dict(W=203, H=117, n=64, seed=8, fps=6),
                            dict(CLI, fps=6, box_size=203, blur_scale=24, threshold=5,
                                 mask_areas=[((20, 10), (60, 40)), ((100, 5), (180, 60), (120, 110), (90, 70))])),
}


def main():
    for name, (clip, kw) in CASES.items():
        frames = synth.make_clip(clip["W"], clip["H"], clip["n"], clip["seed"], fps=clip["fps"])
        out = ref_loader.run_reference(list(frames), **kw)
        fx = {"clip": clip, "kwargs": kw, "params": out["params"],
              "result": [bool(out["result"][0]), out["result"][1], list(out["result"][2])],
              "writes": out["writes"], "trace": out["trace"]}
        path = os.path.join(HERE, name + ".json")
        with open(path, "w") as f:
            json.dump(fx, f, separators=(",", ":"))
        nmov = sum(e["movement"] for e in out["trace"])
        print(f"{name}: {len(out['trace'])} frames, {nmov} with movement, writes={out['writes']}, "
              f"params={out['params']}, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()

"""Seeded synthetic camera clips (numpy only; no cv2) -- SURVEY.md section 8(d).

A clip is a static textured background plus small per-frame sensor noise (stays below the
motion threshold after blurring, so quiet frames have zero contours) plus scripted objects:

  * ``walker``  a bright filled rectangle (W/10 x H/4) translating W/40 px per frame,
  * ``blip``    two small blobs present for 3 frames (exercises the per-contour counter),
  * ``ring``    a ring with a dot in its hole (exercises RETR_EXTERNAL nesting / hole fill),
  * ``hidden``  a blob that moves entirely inside a masked polygon (must be ignored).

Frames are uint8 BGR, HWC, exactly what cv2.VideoCapture.read() hands the reference
(find_motion.py:501).  The same generator feeds the CUDA path, the oracle, the golden
fixture script and bench.py, so every arm sees identical pixels for a given seed.
"""
from __future__ import annotations

import numpy as np

# README.md:36-37 masks (square + triangle) plus a translated pair, SURVEY.md section 8(d) cfg 2
README_MASKS = [((0, 0), (100, 100)), ((0, 0), (0, 100), (100, 0))]
CFG2_MASKS = README_MASKS + [((1500, 200), (1800, 500)), ((300, 700), (300, 1000), (700, 700))]


def static_background(W: int, H: int, rng: np.random.Generator) -> np.ndarray:
    """Blocky low-frequency texture + gradient, values kept inside [24, 200]."""
    cell = max(8, W // 40)
    gh, gw = H // cell + 2, W // cell + 2
    coarse = rng.integers(40, 160, size=(gh, gw, 3), dtype=np.int32)
    up = np.repeat(np.repeat(coarse, cell, axis=0), cell, axis=1)[:H, :W]
    gx = (np.arange(W, dtype=np.int32) * 40 // max(W - 1, 1))[None, :, None]
    gy = (np.arange(H, dtype=np.int32) * 24 // max(H - 1, 1))[:, None, None]
    fine = rng.integers(-8, 9, size=(H, W, 3), dtype=np.int32)
    return np.clip(up + gx - gy + fine, 24, 200).astype(np.int16)


def default_script(W: int, H: int, n_frames: int, fps: int = 30):
    """Events scaled to the clip length: (kind, first_frame, last_frame_exclusive).

    Layout (fractions of the clip): a 3-frame two-blob blip, a blob that lives inside the
    README square mask, a walker episode, a long quiet gap (movement decay runs out and the
    frame cache refills), a ring-with-dot episode, and a quiet tail.
    """
    ev = []
    n = n_frames
    if n >= 12:
        a = max(2, n // 16)
        ev.append(("blip", a, a + 3))
    if n >= 16:
        ev.append(("hidden", n // 8, n // 8 + max(4, n // 6)))
    if n >= 24:
        a = n // 5
        ev.append(("walker", a, min(n, a + max(6, n // 7))))
    if n >= 48:
        a = (n * 5) // 8
        ev.append(("ring", a, min(n, a + max(4, n // 12))))
    return ev


def _paint(frame: np.ndarray, kind: str, k: int, W: int, H: int) -> None:
    """Draw object ``kind`` at the k-th frame of its episode (in place, int16 BGR)."""
    if kind == "walker":
        bw, bh = max(4, W // 10), max(4, H // 4)
        x0 = (W // 8 + k * max(1, W // 40)) % max(1, W - bw)
        y0 = H // 2 - bh // 2
        frame[y0:y0 + bh, x0:x0 + bw] = (235, 240, 245)
    elif kind == "blip":
        s = max(3, W // 64)
        for cx, cy in ((W // 2, H // 5), (W // 2 + 6 * s, H // 5 + 3 * s)):
            frame[cy:cy + s, cx:cx + s] = (250, 250, 250)
    elif kind == "ring":
        R = max(10, H // 6)
        r_in = (R * 3) // 5
        cx, cy = (W * 3) // 4, (H * 2) // 3
        cx = min(cx + k, W - R - 2)
        y, x = np.ogrid[cy - R:cy + R + 1, cx - R:cx + R + 1]
        d2 = (x - cx) ** 2 + (y - cy) ** 2
        sub = frame[cy - R:cy + R + 1, cx - R:cx + R + 1]
        ring = (d2 <= R * R) & (d2 >= r_in * r_in)
        dot = d2 <= max(2, R // 8) ** 2
        sub[ring | dot] = (10, 250, 250)
    elif kind == "hidden":
        # stays inside the README square mask ((0,0),(100,100)) in source coordinates
        s = 12
        x0 = 10 + (3 * k) % 60
        y0 = 10 + (2 * k) % 60
        if x0 + s < W and y0 + s < H:
            frame[y0:y0 + s, x0:x0 + s] = (255, 255, 255)


def make_clip(W: int, H: int, n_frames: int, seed: int, script=None, fps: int = 30,
              noise: int = 3) -> np.ndarray:
    """Return uint8 array (n_frames, H, W, 3), BGR."""
    rng = np.random.default_rng(seed)
    bg = static_background(W, H, rng)
    if script is None:
        script = default_script(W, H, n_frames, fps)
    out = np.empty((n_frames, H, W, 3), dtype=np.uint8)
    for t in range(n_frames):
        f = bg + rng.integers(-noise, noise + 1, size=(H, W, 3), dtype=np.int16)
        for kind, a, b in script:
            if a <= t < b:
                _paint(f, kind, t - a, W, H)
        np.clip(f, 0, 255, out=f)
        out[t] = f.astype(np.uint8)
    return out


def stream_seed(cfg: int, stream_id: int) -> int:
    """seed = 1000*cfg + stream_id."""
    return 1000 * cfg + stream_id

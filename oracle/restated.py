"""TEST INFRASTRUCTURE ONLY -- CPU restatement of find_motion's per-frame hot path.

This is the *oracle* the CUDA path is checked against.  It restates, in numpy integer /
strictly ordered IEEE arithmetic (scipy.ndimage.label for the two labellings), what the
reference's cv2 call chain computes (find_motion/find_motion.py:487-494, 619-700, 549-589).
The pixel arithmetic itself lives in third-party code that is not under /root/reference:
``opencv-python`` (unpinned: requirements.txt:4, setup.py:10) and ``imutils`` (unpinned:
requirements.txt:5).  The behavioural pin is opencv-python-headless 4.13.0.92 on an
AVX2-capable x86-64 host with cv2.setUseOptimized(True).

Pinning: the reference has no tests or golden vectors of its own.
This restatement is pinned by (1) tests/test_oracle_vs_cv2.py, which checks it stage by stage
against cv2 and (test_against_live_reference_loop) against the *real* reference loop
(oracle/ref_loader.py) in the build container, and (2) tests/golden/*.json, traces generated from
the real reference by tests/golden/make_golden.py and committed, which travel to the GPU box
(tests/test_oracle_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this package.  The product (find_motion_b200/) never does.
"""
from __future__ import annotations

import math
from collections import deque

import numpy as np

# --------------------------------------------------------------------------------------
# A.0 derived parameters (find_motion.py:334-335, 406, 422-423, 482-484; imutils.resize)
# --------------------------------------------------------------------------------------


def derive_params(W, H, fps=30, box_size=100, min_box_scale=50, cache_time=2.0, min_time=0.5,
                  blur_scale=20):
    g = int(box_size / blur_scale)
    k = g + 1 if g % 2 == 0 else g
    scale = box_size / W
    return {
        "cache_frames": int(cache_time * fps),            # find_motion.py:334
        "min_movement_frames": int(min_time * fps),       # find_motion.py:335
        "min_area": int(math.pow(box_size / min_box_scale, 2)),   # find_motion.py:406
        "scale": scale,                                   # find_motion.py:422
        "max_area": int((W * H) / 2 * scale),             # find_motion.py:423
        "gaussian": k,                                    # find_motion.py:482-484
        "w": box_size,                                    # imutils.resize(width=box_size)
        "h": int(H * (box_size / float(W))),
    }


# --------------------------------------------------------------------------------------
# A.1 INTER_AREA resize (find_motion.py:492 -> imutils.resize -> cv2.resize INTER_AREA)
# --------------------------------------------------------------------------------------


def area_tab(src: int, dst: int):
    """Per destination index: list of (source index, float32 weight), ascending."""
    scale = 1.0 / (dst / src)          # cv2 computes inv_scale = dsize/ssize, scale = 1/inv_scale
    tabs = []
    for d in range(dst):
        f1 = d * scale
        f2 = f1 + scale
        cell = min(scale, src - f1)
        s1 = math.ceil(f1)
        s2 = min(math.floor(f2), src - 1)
        s1 = min(s1, s2)
        t = []
        if s1 - f1 > 1e-3:
            t.append((s1 - 1, np.float32((s1 - f1) / cell)))
        for s in range(s1, s2):
            t.append((s, np.float32(1.0 / cell)))
        if f2 - s2 > 1e-3:
            t.append((s2, np.float32(min(min(f2 - s2, 1.0), cell) / cell)))
        tabs.append(t)
    return tabs


def _pad_tab(tabs):
    m = max(len(t) for t in tabs)
    idx = np.zeros((len(tabs), m), np.int64)
    wt = np.zeros((len(tabs), m), np.float32)
    for d, t in enumerate(tabs):
        for j, (s, a) in enumerate(t):
            idx[d, j] = s
            wt[d, j] = a
        for j in range(len(t), m):       # zero weight padding: x + 0.0f == x for x >= 0
            idx[d, j] = t[-1][0]
    return idx, wt


def resize_area(img: np.ndarray, w: int, h: int) -> np.ndarray:
    """cv2.resize(img, (w, h), interpolation=INTER_AREA) for u8 HxWx3, downscale/identity."""
    H, W = img.shape[:2]
    if (w, h) == (W, H):
        return img.copy()
    if w > W or h > H:
        raise ValueError("INTER_AREA upscaling is out of scope (SURVEY.md A.1)")
    if W % w == 0 and H % h == 0:
        fx, fy = W // w, H // h
        s = img.reshape(h, fy, w, fx, -1).astype(np.int64).sum(axis=(1, 3))
        if fx == 2 and fy == 2:
            return ((s + 2) >> 2).astype(np.uint8)
        sc = np.float32(1.0 / (fx * fy))
        v = np.rint(s.astype(np.float32) * sc)
        return np.clip(v, 0, 255).astype(np.uint8)
    xi, xw = _pad_tab(area_tab(W, w))
    yi, yw = _pad_tab(area_tab(H, h))
    S = img.astype(np.float32)
    buf = np.zeros((H, w) + img.shape[2:], np.float32)
    for j in range(xi.shape[1]):           # ascending x taps, unfused float32 mul then add
        a = xw[:, j].reshape((1, w) + (1,) * (img.ndim - 2))
        buf = buf + S[:, xi[:, j]] * a
    out = None
    for j in range(yi.shape[1]):           # ascending y taps
        b = yw[:, j].reshape((h, 1) + (1,) * (img.ndim - 2))
        term = buf[yi[:, j]] * b
        out = term if out is None else out + term
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


# --------------------------------------------------------------------------------------
# A.2 gray (find_motion.py:493)
# --------------------------------------------------------------------------------------


def bgr2gray(bgr: np.ndarray) -> np.ndarray:
    b = bgr[..., 0].astype(np.int32)
    g = bgr[..., 1].astype(np.int32)
    r = bgr[..., 2].astype(np.int32)
    return ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15).astype(np.uint8)


# --------------------------------------------------------------------------------------
# A.3 Gaussian blur, sigma 0, u8 (find_motion.py:494)
# --------------------------------------------------------------------------------------

_SMALL = {
    1: [1.0],
    3: [0.25, 0.5, 0.25],
    5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
    7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125],
}


def gauss_coeffs(k: int) -> np.ndarray:
    """8.8 fixed-point taps of cv2.GaussianBlur(u8, (k,k), 0); sum == 256."""
    if k in _SMALL:
        kern = list(_SMALL[k])
    else:
        sigma = 0.3 * ((k - 1) * 0.5 - 1) + 0.8
        scale2x = -0.5 / (sigma * sigma)
        t = [math.exp(scale2x * (i - (k - 1) * 0.5) ** 2) for i in range(k)]
        s = 0.0
        for v in t:
            s += v
        inv = 1.0 / s
        kern = [v * inv for v in t]
    c = [0] * k
    err = 0.0
    tot = 0
    for i in range(k // 2):
        adj = kern[i] * 256.0 + err
        v = int(np.rint(adj))
        err = adj - v
        c[i] = c[k - 1 - i] = v
        tot += 2 * v
    c[k // 2] = 256 - tot
    return np.array(c, np.int64)


def reflect101(i: np.ndarray, n: int) -> np.ndarray:
    """BORDER_REFLECT_101 index map, valid for any offset (repeated reflection)."""
    if n == 1:
        return np.zeros_like(i)
    p = 2 * (n - 1)
    i = np.mod(i, p)
    return np.where(i >= n, p - i, i)


def gaussian_blur(gray: np.ndarray, k: int) -> np.ndarray:
    h, w = gray.shape
    c = gauss_coeffs(k)
    r = k // 2
    g = gray.astype(np.int64)
    xs = reflect101(np.arange(-r, w + r), w)
    gp = g[:, xs]
    hor = np.zeros((h, w), np.int64)
    for i in range(k):
        if c[i]:
            hor += c[i] * gp[:, i:i + w]
    ys = reflect101(np.arange(-r, h + r), h)
    hp = hor[ys]
    ver = np.zeros((h, w), np.int64)
    for j in range(k):
        if c[j]:
            ver += c[j] * hp[j:j + h]
    return ((ver + 32768) >> 16).astype(np.uint8)


# --------------------------------------------------------------------------------------
# A.5 masks (find_motion.py:611-635): cv2.rectangle FILLED / cv2.fillConvexPoly
# --------------------------------------------------------------------------------------


def scale_area(area, scale):
    return [(int(a[0] * scale), int(a[1] * scale)) for a in area]     # find_motion.py:616


def _cdiv(a: int, b: int) -> int:
    """C integer division (truncation toward zero)."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def _clip_line(w, h, p1, p2):
    """cv2.clipLine on the rectangle [0,w) x [0,h); returns (ok, p1, p2)."""
    x1, y1 = p1
    x2, y2 = p2
    right, bottom = w - 1, h - 1
    if w <= 0 or h <= 0:
        return False, p1, p2
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += _cdiv((a - y1) * (x2 - x1), (y2 - y1))
            y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += _cdiv((a - y2) * (x2 - x1), (y2 - y1))
            y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += _cdiv((a - x1) * (y2 - y1), (x2 - x1))
                x1 = a
                c1 = 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += _cdiv((a - x2) * (y2 - y1), (x2 - x1))
                x2 = a
                c2 = 0
    return (c1 | c2) == 0, (x1, y1), (x2, y2)


def _line8(mask, p1, p2):
    """8-connected Bresenham as cv2's LineIterator draws it (left to right)."""
    h, w = mask.shape
    ok, p1, p2 = _clip_line(w, h, p1, p2)
    if not ok:
        return
    x1, y1 = p1
    x2, y2 = p2
    dx, dy = x2 - x1, y2 - y1
    # LineIterator(leftToRight=true): start from the point with the smaller x
    if dx < 0:
        x1, y1, x2, y2 = x2, y2, x1, y1
        dx, dy = -dx, -dy
    sy = 1 if dy >= 0 else -1
    ady = abs(dy)
    if dx >= ady:            # x major
        major, minor = dx, ady
        err = major - 2 * minor
        x, y = x1, y1
        for _ in range(major + 1):
            mask[y, x] = True
            if err < 0:
                y += sy
                err += 2 * major - 2 * minor
            else:
                err -= 2 * minor
            x += 1
    else:                    # y major
        major, minor = ady, dx
        err = major - 2 * minor
        x, y = x1, y1
        for _ in range(major + 1):
            mask[y, x] = True
            if err < 0:
                x += 1
                err += 2 * major - 2 * minor
            else:
                err -= 2 * minor
            y += sy


def fill_convex_poly(mask: np.ndarray, pts) -> None:
    """cv2.fillConvexPoly(img, pts, 0) footprint (shift 0, 8-connected) -> mask |= footprint."""
    h, w = mask.shape
    n = len(pts)
    XY_SHIFT, XY_ONE = 16, 1 << 16
    ys_ = [p[1] for p in pts]
    ymin, ymax = min(ys_), max(ys_)
    imin = ys_.index(ymin)
    xs_ = [p[0] for p in pts]
    xmin, xmax = min(xs_), max(xs_)
    # outline
    p0 = pts[n - 1]
    for i in range(n):
        p = pts[i]
        _line8(mask, p0, p)
        p0 = p
    if n < 3 or xmax < 0 or ymax < 0 or xmin >= w or ymin >= h:
        return
    ymax = min(ymax, h - 1)
    edge = [{"idx": imin, "di": 1, "ye": ymin, "x": -XY_ONE, "dx": 0},
            {"idx": imin, "di": n - 1, "ye": ymin, "x": -XY_ONE, "dx": 0}]
    edges = n
    y = ymin
    left, right = 0, 1
    while True:
        for i in range(2):
            e = edge[i]
            if y >= e["ye"]:
                idx0, di = e["idx"], e["di"]
                idx = idx0 + di
                if idx >= n:
                    idx -= n
                while True:             # C: for (; edges-- > 0; )
                    go = edges > 0
                    edges -= 1
                    if not go:
                        break
                    ty = pts[idx][1]
                    if ty > y:
                        xs = pts[idx0][0]
                        xe = pts[idx][0]
                        e["ye"] = ty
                        num = ((xe - xs) << XY_SHIFT) * 2 + (ty - y)
                        e["dx"] = _cdiv(num, 2 * (ty - y))
                        e["x"] = xs << XY_SHIFT
                        e["idx"] = idx
                        break
                    idx0 = idx
                    idx += di
                    if idx >= n:
                        idx -= n
        if edges < 0:
            break
        if y >= 0:
            l, r = 0, 1
            if edge[0]["x"] > edge[1]["x"]:
                l, r = 1, 0
            xx1 = (edge[l]["x"] + (XY_ONE >> 1)) >> XY_SHIFT
            xx2 = (edge[r]["x"] + (XY_ONE >> 1)) >> XY_SHIFT
            if xx2 >= 0 and xx1 < w:
                xx1 = max(xx1, 0)
                xx2 = min(xx2, w - 1)
                if xx2 >= xx1:
                    mask[y, xx1:xx2 + 1] = True
        edge[0]["x"] += edge[0]["dx"]
        edge[1]["x"] += edge[1]["dx"]
        y += 1
        if y > ymax:
            break


def rasterise_masks(w: int, h: int, mask_areas, scale: float) -> np.ndarray:
    """Boolean (h, w) plane: True where find_motion.py:619-635 paints BLACK into blur."""
    m = np.zeros((h, w), bool)
    for area in mask_areas or []:
        pts = scale_area(area, scale)
        if len(pts) == 2:          # cv2.rectangle FILLED: inclusive corners, clipped
            (xa, ya), (xb, yb) = pts
            x0, x1 = max(min(xa, xb), 0), min(max(xa, xb), w - 1)
            y0, y1 = max(min(ya, yb), 0), min(max(ya, yb), h - 1)
            if x1 >= x0 and y1 >= y0:
                m[y0:y1 + 1, x0:x1 + 1] = True
        else:
            fill_convex_poly(m, pts)
    return m


# --------------------------------------------------------------------------------------
# A.6 / A.7 background, diff, threshold (find_motion.py:246-257, 651-659)
# --------------------------------------------------------------------------------------


def _two_prod(a, b):
    """Dekker/Veltkamp product: a*b == p + e exactly (float64 arrays, no overflow here)."""
    p = a * b
    split = 134217729.0      # 2**27 + 1
    ca = split * a
    ah = ca - (ca - a)
    al = a - ah
    cb = split * b
    bh = cb - (cb - b)
    bl = b - bh
    e = ((ah * bh - p) + ah * bl + al * bh) + al * bl
    return p, e


def _two_sum(a, b):
    """Knuth two-sum: a + b == s + t exactly."""
    s = a + b
    bb = s - a
    t = (a - (s - bb)) + (b - bb)
    return s, t


def fma_f64(a, b, c):
    """Vectorised correctly rounded a*b + c (one rounding), without math.fma (Python 3.12).

    a*b = p + e and p + c = s + t exactly; corr = rn(t + e) is off by at most 2**-53*|corr|,
    so rn(s + corr) is the correctly rounded result unless s + corr sits within that sliver
    of a rounding boundary.  Those elements (and exact ties) are redone with Fractions.
    """
    from fractions import Fraction

    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(np.broadcast_to(np.asarray(b, np.float64), a.shape))
    c = np.ascontiguousarray(c, np.float64)
    p, e = _two_prod(a, b)
    s, t = _two_sum(p, c)
    corr = t + e
    r, u = _two_sum(s, corr)
    sp = np.spacing(np.abs(r))
    au = np.abs(u)
    m, _ = np.frexp(r)
    doubtful = (au >= sp * (0.5 - 2.0 ** -30)) | ((np.abs(m) == 0.5) & (au >= sp * (0.25 - 2.0 ** -30)))
    if np.any(doubtful):
        ra, rb, rc, rr = a.ravel(), b.ravel(), c.ravel(), r.ravel()
        for i in np.nonzero(doubtful.ravel())[0]:
            rr[i] = float(Fraction(float(ra[i])) * Fraction(float(rb[i])) + Fraction(float(rc[i])))
    return r


def accumulate_weighted(bg: np.ndarray, src: np.ndarray, alpha: float) -> np.ndarray:
    """cv2.accumulateWeighted(src u8, bg f64, alpha): AVX2 body + contracted scalar tail."""
    a = np.float64(alpha)
    b = np.float64(1.0) - a
    flat = bg.reshape(-1)
    s = src.reshape(-1).astype(np.float64)
    n = flat.size
    nb = n - (n % 16)
    out = np.empty_like(flat)
    out[:nb] = fma_f64(flat[:nb], b, s[:nb] * a)        # fma(bg, beta, rn(src*alpha))
    if nb < n:
        out[nb:] = fma_f64(s[nb:], a, flat[nb:] * b)    # fma(src, alpha, rn(bg*beta))
    return out.reshape(bg.shape)


def bg_to_u8(bg: np.ndarray) -> np.ndarray:
    """cv2.convertScaleAbs on float64: double -> float32 -> |.| -> rne -> saturate."""
    f = np.abs(bg.astype(np.float32))
    return np.clip(np.rint(f), 0, 255).astype(np.uint8)


def diff_threshold(blur: np.ndarray, bg: np.ndarray, threshold: int) -> np.ndarray:
    d = np.abs(blur.astype(np.int16) - bg_to_u8(bg).astype(np.int16))
    return np.where(d > threshold, 255, 0).astype(np.uint8)


# --------------------------------------------------------------------------------------
# A.8 dilate + external contours (find_motion.py:260-276, 679, 792)
# --------------------------------------------------------------------------------------


def dilate5(t: np.ndarray) -> np.ndarray:
    """cv2.dilate(kernel=None, iterations=2) == 5x5 max, out-of-image ignored."""
    h, w = t.shape
    p = np.zeros((h + 4, w + 4), t.dtype)
    p[2:-2, 2:-2] = t
    out = np.zeros_like(t)
    for dy in range(5):
        for dx in range(5):
            np.maximum(out, p[dy:dy + h, dx:dx + w], out=out)
    return out


def external_components(binary: np.ndarray):
    """[(area_x2, (x, y, w, h))...] of cv2.findContours(RETR_EXTERNAL) contours.

    area_x2 is 2*cv2.contourArea (an integer); boxes are cv2.boundingRect.  Sorted.
    """
    from scipy import ndimage

    fg = binary != 0
    h, w = fg.shape
    if not fg.any():
        return []
    pad = np.zeros((h + 2, w + 2), bool)
    pad[1:-1, 1:-1] = fg
    four = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], bool)
    lab, _ = ndimage.label(~pad, structure=four)
    outside = lab == lab[0, 0]
    F = ~outside                                   # foreground with holes filled (padded)
    lab8, n = ndimage.label(F, structure=np.ones((3, 3), bool))
    if n == 0:
        return []
    Fi = F.astype(np.int32)
    q = Fi[:-1, :-1] + Fi[:-1, 1:] + Fi[1:, :-1] + Fi[1:, 1:]       # 2x2 windows
    # a window with >= 3 set pixels lies inside one component: take the max label in it
    l = np.maximum(np.maximum(lab8[:-1, :-1], lab8[:-1, 1:]), np.maximum(lab8[1:, :-1], lab8[1:, 1:]))
    q4 = np.bincount(l[q == 4], minlength=n + 1)
    q3 = np.bincount(l[q == 3], minlength=n + 1)
    out = []
    objs = ndimage.find_objects(lab8)
    for i in range(1, n + 1):
        sy, sx = objs[i - 1]
        x0, y0 = sx.start - 1, sy.start - 1
        out.append((int(2 * q4[i] + q3[i]), (int(x0), int(y0), int(sx.stop - sx.start), int(sy.stop - sy.start))))
    out.sort()
    return out


# --------------------------------------------------------------------------------------
# A.9 decision state machine (find_motion.py:665-700, 549-589)
# --------------------------------------------------------------------------------------


class Decision:
    """Counters + bounded frame cache of one stream.  step() returns the per-frame record."""

    def __init__(self, min_area, max_area, cache_frames, min_movement_frames):
        self.min_area = min_area
        self.max_area = max_area
        self.cache_frames = cache_frames
        self.min_movement_frames = min_movement_frames
        self.counter = 0
        self.decay = 0
        self.cache_len = 0
        self.wrote_frames = False

    def step(self, areas_x2):
        movement = False
        if self.decay > 0:
            self.decay -= 1
        for a2 in areas_x2:
            area = a2 / 2.0
            if self.max_area < area < self.min_area:
                continue
            self.counter += 1
            movement = True
        if not movement:
            self.counter = 0
        n_flush = 0
        wrote = False
        if self.counter >= self.min_movement_frames or self.decay > 0:
            if movement:
                self.decay = self.cache_frames
                n_flush = self.cache_len
                self.cache_len = 0
            wrote = True
            self.wrote_frames = True
        else:
            # deque(maxlen=cache_frames).append: oldest silently dropped (find_motion.py:415, 588)
            self.cache_len = min(self.cache_len + 1, self.cache_frames)
        if n_flush:
            self.wrote_frames = True
        return {"movement": movement, "counter": self.counter, "decay": self.decay,
                "cache_len": self.cache_len, "wrote": wrote, "n_flush": n_flush}


# --------------------------------------------------------------------------------------
# whole-stream driver (find_motion.py:852-904)
# --------------------------------------------------------------------------------------


class StreamOracle:
    """One stream's state: background plane + Decision.  process(frame) -> per-frame dict."""

    def __init__(self, W, H, fps=30, box_size=100, min_box_scale=50, cache_time=2.0, min_time=0.5,
                 threshold=7, avg=0.1, blur_scale=20, mask_areas=None):
        self.W, self.H = W, H
        self.p = derive_params(W, H, fps, box_size, min_box_scale, cache_time, min_time, blur_scale)
        self.w, self.h, self.k = self.p["w"], self.p["h"], self.p["gaussian"]
        self.threshold = threshold
        self.avg = avg
        self.mask = rasterise_masks(self.w, self.h, mask_areas, self.p["scale"])
        self.bg = None
        self.dec = Decision(self.p["min_area"], self.p["max_area"], self.p["cache_frames"],
                            self.p["min_movement_frames"])

    def process(self, frame: np.ndarray, keep_planes=False):
        small = resize_area(frame, self.w, self.h)
        gray = bgr2gray(small)
        blur = gaussian_blur(gray, self.k)
        blur[self.mask] = 0
        if self.bg is None:
            self.bg = blur.astype(np.float64)
        thresh = diff_threshold(blur, self.bg, self.threshold)
        self.bg = accumulate_weighted(self.bg, blur, self.avg)
        dil = dilate5(thresh)
        comps = external_components(dil)
        rec = self.dec.step([a for a, _ in comps])
        rec.update(areas=sorted(a / 2.0 for a, _ in comps), boxes=sorted(b for _, b in comps))
        if keep_planes:
            rec["planes"] = {"gray": gray, "blur": blur, "thresh": dil, "bg": self.bg.copy()}
        return rec

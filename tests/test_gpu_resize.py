"""Default (decimating) mode front end: the row-per-lane INTER_AREA + gray kernel (csrc/k_resize_rows.cu) against cv2's own
resize + cvtColor -- the calls of blur_frame (find_motion.py:487-493) -- and against the warp-per-destination-row kernels
it replaces, for exact equality, over the ratios and alignments that select its code paths (tap groups per column, byte
phase of a column's first tap, several passes per column, bands of different heights, ragged batches)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GEOMS = [(1920, 1080, 100),      # the reference's CLI default on 1080p: 19.2 : 1, 20 taps per column
         (1280, 720, 100),       # 12.8 : 1
         (640, 480, 100),        # 6.4 : 1 (golden cfg1's geometry), 24 destination rows per band
         (3840, 2160, 100),      # 38.4 : 1: two passes per column, 4 destination rows per band
         (1920, 1080, 333),      # 5.77 : 1, plane wider than one segment row
         (1920, 1080, 1000),     # 1.92 : 1: falls back (fewer than 4 taps) -- the old kernels stay exact
         (2560, 1440, 150),
         (1936, 1096, 123)]      # rows of 5808 bytes: 16-byte aligned but not a multiple of 128


@pytest.mark.parametrize("W,H,box", GEOMS)
def test_resize_gray_matches_cv2(W, H, box):
    cv2 = pytest.importorskip("cv2")
    import torch
    from find_motion_b200.engine import MotionEngine
    rng = np.random.default_rng(W + box)
    S, T = 2, 3
    frames = rng.integers(0, 256, size=(S, T, H, W, 3), dtype=np.uint8)
    frames[0, 1] = 255                                   # saturated frame: every chain ends at exactly 255
    frames[1, 0, :, : W // 2] = 0
    h = int(H * (box / float(W)))
    dev = torch.from_numpy(frames).cuda()
    got = {}
    for no_rows in (False, True):
        with MotionEngine(W, H, n_streams=S, max_frames=T, box_size=box, blur_scale=20, keep_planes=True,
                          no_rows=no_rows) as eng:
            eng.process(dev, n_valid=[T, T - 1])
            got[no_rows] = [[eng.planes(s, t, gray=True, blur=False, thresh=False, bg=False)["gray"]
                             for t in range(T - (s == 1))] for s in range(S)]
    for s in range(S):
        for t in range(T - (s == 1)):
            want = cv2.cvtColor(cv2.resize(frames[s, t], (box, h), interpolation=cv2.INTER_AREA), cv2.COLOR_BGR2GRAY)
            assert got[False][s][t].shape == want.shape
            assert (got[False][s][t] == want).all(), ("rows kernel", s, t, int((got[False][s][t] != want).sum()))
            assert (got[True][s][t] == want).all(), ("warp kernel", s, t, int((got[True][s][t] != want).sum()))


def test_resize_rows_unaligned_frames_fall_back():
    """Frames whose stream / frame strides are not multiples of 16 bytes cannot go through TMA: same result from the old path."""
    cv2 = pytest.importorskip("cv2")
    import torch
    from find_motion_b200.engine import MotionEngine
    W, H, box = 636, 476, 100                              # 1908-byte rows (not a multiple of 16)
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, size=(1, 2, H, W, 3), dtype=np.uint8)
    h = int(H * (box / float(W)))
    with MotionEngine(W, H, n_streams=1, max_frames=2, box_size=box, keep_planes=True) as eng:
        eng.process(torch.from_numpy(frames).cuda())
        for t in range(2):
            want = cv2.cvtColor(cv2.resize(frames[0, t], (box, h), interpolation=cv2.INTER_AREA), cv2.COLOR_BGR2GRAY)
            assert (eng.planes(0, t, gray=True, blur=False, thresh=False, bg=False)["gray"] == want).all()

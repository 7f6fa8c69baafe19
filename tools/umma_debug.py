"""Development aid: where does the tcgen05 Gaussian differ from the oracle?  Prints the 32-row x 16-column blocks of the
blur plane that mismatch, per frame, and a few value pairs.  usage: python tools/umma_debug.py W H blur_scale apron(0/1)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from find_motion_b200 import synth
from find_motion_b200.engine import MotionEngine
from oracle import restated as R

W, H, bs, apron = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
n = T = 4
kw = dict(fps=6, box_size=W, blur_scale=bs, threshold=5, avg=0.2, min_time=0.3, cache_time=0.6)
clip = synth.make_clip(W, H, n, seed=7, fps=6)[None]
orc = R.StreamOracle(W, H, **kw)
dev = torch.from_numpy(clip).cuda()
for rep in range(2):
    with MotionEngine(W, H, n_streams=1, max_frames=T, keep_planes=True, no_fused=True, umma=True, umma_apron=bool(apron), **kw) as eng:
        eng.process(dev)
        orc = R.StreamOracle(W, H, **kw)
        for t in range(n):
            rec = orc.process(clip[0, t], keep_planes=True)
            pl = eng.planes(0, t)
            bad = pl["blur"] != rec["planes"]["blur"]
            blocks = sorted({(int(y) // 32, int(x) // 16) for y, x in np.argwhere(bad)})
            print(f"rep {rep} frame {t}: {int(bad.sum())} bad pixels, blocks (row/32, col/16): {blocks[:40]}")
            ys, xs = np.nonzero(bad)
            if len(ys):
                rows = sorted(set(ys.tolist()))
                print("   rows", rows[:8], "...", rows[-4:], "cols", sorted(set(xs.tolist()))[:20])
                for y, x in list(zip(ys, xs))[:6]:
                    print("   ", y, x, "got", pl["blur"][y, x], "want", rec["planes"]["blur"][y, x])

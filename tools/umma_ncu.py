"""One call of the tcgen05 blur kernel on the bench workload, for `ncu -k regex:k_umma`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from find_motion_b200 import synth
from find_motion_b200.engine import MotionEngine
W, H, S, T = 1920, 1080, 8, 16
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 20
apron = len(sys.argv) > 2 and sys.argv[2] == "apron"
kw = dict(fps=30, box_size=W, blur_scale=bs, threshold=12, avg=0.1, min_time=0.5, cache_time=1.0, mask_areas=synth.CFG2_MASKS)
clips = torch.stack([torch.from_numpy(synth.make_clip(W, H, T, seed=2000 + s)) for s in range(2)]).cuda()
frames = clips.repeat(S // 2, 1, 1, 1, 1).contiguous()
with MotionEngine(W, H, n_streams=S, max_frames=T, no_fused=True, umma=True, umma_apron=apron, **kw) as eng:
    for i in range(3):
        eng.process(frames, sync=False)
    torch.cuda.synchronize()
print("done")

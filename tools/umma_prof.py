"""Development aid: per-phase cycle counts of the tcgen05 blur kernel (apron variant), FM_UMMA_PROF=1."""
import ctypes, os, sys
os.environ["FM_UMMA_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from find_motion_b200 import _lib, synth
from find_motion_b200.engine import MotionEngine

W, H, S, T = 1920, 1080, 8, 16
lib = _lib.load()
for bs in (20, 384):
    kw = dict(fps=30, box_size=W, blur_scale=bs, threshold=12, avg=0.1, min_time=0.5, cache_time=1.0, mask_areas=synth.CFG2_MASKS)
    clip = torch.from_numpy(synth.make_clip(W, H, T, seed=1)).cuda()
    frames = clip[None].expand(S, T, H, W, 3).contiguous()
    with MotionEngine(W, H, n_streams=S, max_frames=T, no_fused=True, umma=True, umma_apron=True, **kw) as eng:
        for i in range(3):
            eng.process(frames, sync=False)
        torch.cuda.synchronize()
        lib.fm_umma_prof_dump()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10):
            eng.process(frames, sync=False)
        e1.record(); torch.cuda.synchronize()
        print("blur_scale", bs, "k", eng.info["gaussian"], "ms/call", e0.elapsed_time(e1) / 10, file=sys.stderr)
        lib.fm_umma_prof_dump()

"""find_motion_b200: B200-native (sm_100a) per-frame motion-detection hot path of
dmiruke/find_motion behind the reference's own Python surface.  See DESIGN.md."""
__all__ = ["MotionEngine", "VideoMotion", "run_vid", "synth"]


def __getattr__(name):
    if name == "MotionEngine":
        from .engine import MotionEngine
        return MotionEngine
    if name in ("VideoMotion", "run_vid"):
        from . import video_motion
        return getattr(video_motion, name)
    raise AttributeError(name)

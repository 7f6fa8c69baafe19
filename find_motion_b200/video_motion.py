"""Drop-in mirror of the reference's per-stream surface: VideoMotion and run_vid
(find_motion/find_motion.py:299-380, 852-904, 1021-1037), with the per-frame cv2 call chain
replaced by the B200 path (libfmgpu.so through MotionEngine).

Same constructor keywords, same derived parameters, same `(wrote_frames, err_msg, seen_objects)`
result, same output files: decode (`cap.read`) and encode (`outfile.write(frame.raw)`) stay on
the host exactly as in the reference; frames are batched `chunk` at a time, the GPU returns the
per-frame decisions, and the adapter replays decide_output's cache/flush/write actions in frame
order on the raw frames it kept.  Display (`show`), Haar cascades and YOLO are outside the hot
path and are ignored with a warning.
"""
from __future__ import annotations

import logging
import os
import typing
from collections import deque

import numpy as np

from .engine import MotionEngine

log = logging.getLogger("find_motion")


class VideoError(Exception):
    """find_motion.py:173-178"""


class VideoMotion(object):
    def __init__(self, filename: typing.Union[str, int, typing.Any] = None, outdir: str = '', fps: int = 30,
                 box_size: int = 100, min_box_scale: int = 50, cache_time: float = 2.0, min_time: float = 0.5,
                 threshold: int = 7, avg: float = 0.1, blur_scale: int = 20,
                 mask_areas: list = None, show: bool = False,
                 codec: str = 'MJPG', log_level: int = logging.INFO,
                 mem: bool = False, cleanup: bool = False,
                 multiprocess: bool = False,
                 cascades: typing.List[str] = None,
                 yolo_tiny: bool = False, *,
                 device: int = 0, chunk: int = 16) -> None:
        self.filename = filename
        if self.filename is None:
            raise Exception('Filename required')                      # find_motion.py:311
        self.log = logging.getLogger('find_motion.VideoMotion')
        self.log.setLevel(log_level)
        self.multiprocess = multiprocess
        self.outfile = None
        self.outfiles = 0
        self.outfile_name = ''
        self.outdir = os.path.normpath(outdir) if outdir != '' else ''
        self.fps = fps
        self.box_size = box_size
        self.min_box_scale = min_box_scale
        self.gaussian_scale = blur_scale
        self.cache_frames = int(cache_time * fps)                     # :334
        self.min_movement_frames = int(min_time * fps)                # :335
        self.delta_thresh = threshold
        self.avg = avg
        self.mask_areas = mask_areas if mask_areas is not None else []
        self.show = show
        self.codec = codec
        self.debug = log_level == logging.DEBUG
        self.mem = mem
        self.cleanup_flag = cleanup
        if show or cascades or yolo_tiny:
            self.log.warning('show / cascades / yolo_tiny are outside the GPU hot path and are ignored')
        self._tuning = dict(fps=fps, box_size=box_size, min_box_scale=min_box_scale, cache_time=cache_time,
                            min_time=min_time, threshold=threshold, avg=avg, blur_scale=blur_scale)
        self.device = device
        self.chunk = max(1, int(chunk))

        self.amount_of_frames = -1
        self.frame_width = -1
        self.frame_height = -1
        self.scale = -1.0
        self.frame_cache: typing.Deque[np.ndarray] = deque()
        self.wrote_frames: typing.Optional[bool] = False
        self.err_msg = ''
        self.movement = False
        self.movement_decay = 0
        self.movement_counter = 0
        self.seen_objects: typing.Set[str] = set()
        self.frames_read = 0
        self.frames_written = 0
        self.engine: typing.Optional[MotionEngine] = None
        self.loaded = self._load_video()

    # -- I/O (host side, as in the reference) ---------------------------------------------------
    def _open_capture(self):
        if hasattr(self.filename, 'read') and hasattr(self.filename, 'get'):
            return self.filename                                     # capture-like object (tests, cameras)
        import cv2
        return cv2.VideoCapture(self.filename)                        # :413

    def _get_video_info(self) -> None:
        import cv2
        self.amount_of_frames = int(self.cap.get(cv2.CAP_PROP_FRAME_COUNT))       # :430-432
        self.frame_width = int(self.cap.get(cv2.CAP_PROP_FRAME_WIDTH))
        self.frame_height = int(self.cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
        if self.frame_width == 0 or self.frame_height == 0:
            broken = 'width' if self.frame_width == 0 else 'height'
            raise VideoError("Video info malformed - {} is 0: {}".format(broken, self.filename))

    def _load_video(self) -> bool:
        self.cap = self._open_capture()
        self.frame_cache = deque(maxlen=self.cache_frames)            # :415
        try:
            self._get_video_info()
        except VideoError as e:
            self.log.error(str(e))
            return False
        self.engine = MotionEngine(self.frame_width, self.frame_height, n_streams=1, max_frames=self.chunk,
                                   device=self.device, mask_areas=self.mask_areas, **self._tuning)
        inf = self.engine.info
        self.scale = inf["scale"]
        self.max_area = inf["max_area"]
        self.min_area = inf["min_area"]
        self.gaussian = (inf["gaussian"], inf["gaussian"])
        return True

    def _make_outfile(self) -> None:                                   # :443-475
        import cv2
        self.outfiles += 1
        if self.outfiles > 1 and self.outfile is not None:
            self.outfile.release()
        outname = str(self.filename) + '_' + str(self.outfiles)
        if self.outdir == '':
            self.outfile_name = outname + '_motion.avi'
        else:
            self.outfile_name = os.path.join(self.outdir, os.path.basename(outname)) + '_motion.avi'
        self.outfile = cv2.VideoWriter(self.outfile_name, cv2.VideoWriter_fourcc(*self.codec), self.fps,
                                       (self.frame_width, self.frame_height))

    def output_raw_frame(self, frame: np.ndarray = None) -> None:      # :533-546 (and :509-530 without show)
        if not self.wrote_frames:
            self._make_outfile()
            self.wrote_frames = True
        try:
            self.outfile.write(frame)
        except Exception as e:
            self.log.warning('Having to create output file due to exception: {}'.format(e))
            self._make_outfile()
            self.outfile.write(frame)
        self.frames_written += 1

    def is_open(self) -> bool:
        return self.cap.isOpened()

    def cleanup(self) -> None:                                         # :907-926
        if getattr(self, 'cap', None) is not None and not hasattr(self.filename, 'read'):
            self.cap.release()
        if self.outfile is not None:
            self.outfile.release()
        if self.engine is not None:
            self.engine.close()
            self.engine = None

    # -- main loop ----------------------------------------------------------------------------------
    def _replay(self, raws, stats) -> None:
        """decide_output's actions on the raw frames, in frame order (find_motion.py:549-589)."""
        for raw, st in zip(raws, stats):
            self.movement = bool(st["movement"])
            self.movement_counter = int(st["movement_counter"])
            self.movement_decay = int(st["movement_decay"])
            if st["wrote"]:
                if st["n_flush"]:
                    assert int(st["n_flush"]) == len(self.frame_cache), "frame cache out of step with the device"
                    for cached in self.frame_cache:
                        self.output_raw_frame(cached)
                    self.frame_cache.clear()
                self.output_raw_frame(raw)
            else:
                self.frame_cache.append(raw)
            assert len(self.frame_cache) == int(st["cache_len"]), "frame cache out of step with the device"

    def find_motion(self) -> tuple:
        """Main loop (find_motion.py:852-904): returns (wrote_frames, err_msg, seen_objects)."""
        batch = np.empty((1, self.chunk, self.frame_height, self.frame_width, 3), np.uint8)
        while self.is_open():
            raws = []
            while len(raws) < self.chunk:
                ret, frame = self.cap.read()
                if not ret:
                    break
                if frame.shape != (self.frame_height, self.frame_width, 3) or frame.dtype != np.uint8:
                    raise VideoError('frame geometry changed mid-stream: {}'.format(frame.shape))
                batch[0, len(raws)] = frame
                raws.append(frame)
            if not raws:
                break
            self.frames_read += len(raws)
            stats = self.engine.process_host(batch[:, :len(raws)])
            self._replay(raws, stats[0])
            if len(raws) < self.chunk:
                break
        self.cleanup()
        return self.wrote_frames, self.err_msg, tuple(self.seen_objects)


def run_vid(filename: typing.Union[str, int], **kwargs) -> tuple:
    """find_motion.py:1021-1037: the job `run()` binds with functools.partial."""
    seen_objects = None
    try:
        vid = VideoMotion(filename=filename, **kwargs)
        if vid.loaded:
            wrote_frames, err_msg, seen_objects = vid.find_motion()
        else:
            wrote_frames = None
            seen_objects = None
            err_msg = 'Video did not load successfully'
    except Exception as e:
        err_msg = 'Error processing video {}: {}'.format(filename, e)
        wrote_frames = None
    return (wrote_frames, filename, err_msg, seen_objects)

"""Output side of the frame loop (reference MAIN:714-731): JPEG files + optional MP4.

The reference encodes and writes two JPEGs per frame with imageio on the render thread, between two frames' kernels.
Here the frames arrive as uint8 [H,W,3] views of pinned host memory (sequence.FrameSink) as soon as their device->host
copy has landed; a small thread pool encodes them (Pillow -- the encoder imageio.imwrite uses, default quality -- which
releases the GIL) while the device renders the following frames.  File names follow the reference: ``test_%06d.jpg``
indexed by the GLOBAL frame number, so frame-sharded ranks (sequence.shard_frames) write disjoint files of one directory.

The optional video (MAIN:727-731 writes ``<expname>.mp4`` at 25 fps through imageio-ffmpeg, absent in this image) is
written with OpenCV's ffmpeg backend (``mp4v``) in frame order at ``close()``; it raises if the codec is unavailable.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np


def _save_jpeg(path, frame):
    from PIL import Image
    Image.fromarray(frame).save(path)         # Pillow defaults (quality 75), as imageio.imwrite(path, rgb8)


class FrameWriter:
    """write(i, frame_u8[H,W,3]) -> <directory>/test_%06d.jpg on a worker pool; close() joins and returns the paths."""

    def __init__(self, directory, workers=4, video=None, fps=25, pattern='test_{:06d}.jpg'):
        os.makedirs(directory, exist_ok=True)
        self.directory, self.pattern = directory, pattern
        self.pool = ThreadPoolExecutor(max_workers=workers)
        self.pending = {}
        self.video, self.fps = video, fps
        self.kept = {} if video else None

    def path(self, i):
        return os.path.join(self.directory, self.pattern.format(i))

    def write(self, i, frame):
        frame = np.asarray(frame)
        if frame.dtype != np.uint8 or frame.ndim != 3 or frame.shape[2] != 3:
            raise ValueError('FrameWriter.write wants uint8 [H,W,3], got %s %s' % (frame.dtype, frame.shape))
        self.pending[i] = self.pool.submit(_save_jpeg, self.path(i), frame)
        if self.kept is not None:
            self.kept[i] = frame

    def close(self):
        paths = []
        for i in sorted(self.pending):
            self.pending[i].result()          # re-raises an encoder / filesystem error of the worker
            paths.append(self.path(i))
        self.pool.shutdown()
        if self.video and self.kept:
            write_video(self.video, [self.kept[i] for i in sorted(self.kept)], self.fps)
        return paths


def write_video(path, frames, fps=25):
    """frames: uint8 RGB [H,W,3] each -> an MP4 (`mp4v`) at `fps`."""
    import cv2
    h, w = frames[0].shape[:2]
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*'mp4v'), float(fps), (w, h))
    if not vw.isOpened():
        raise RuntimeError('cannot open an mp4v VideoWriter for %s (OpenCV built without an ffmpeg encoder?)' % path)
    try:
        for f in frames:
            vw.write(np.ascontiguousarray(f[..., ::-1]))      # OpenCV wants BGR
    finally:
        vw.release()
    return path

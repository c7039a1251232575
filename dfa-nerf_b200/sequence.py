"""Sequence-level frame loop: the caller of the hot path (MAIN:624-733).

The reference renders a driven sequence with one Python iteration per frame: latents, two ``get_rays``, a chunk
loop of ceil(HW/2048) network cascades, ``.cpu().numpy()`` + ``to8b`` + JPEG write -- each frame's device->host copy
and file I/O serialised with the next frame's compute.  Here a frame is ~10 kernel launches, so the loop is
restructured around the device:

* per-frame poses and latents are uploaded once as tables (``[N,3,4]``, ``[N,dim]``);
* the rendered frame is quantised on the device (``dfn_to8b``, HELP:17) and copied to pinned host memory as
  3 bytes/pixel on a side stream, double-buffered, so frame i's copy overlaps frame i+1's render;
* with several GPUs the FRAMES are sharded (contiguous blocks per rank): no collective inside the loop at all,
  one gather of the uint8 frames at the end (BASELINE.json configs[4]: 300-frame sequence on 8 GPUs).

CUDA graphs are deliberately not used: a frame is ten launches of multi-millisecond kernels (launch overhead < 0.1 %).
"""
import torch
import torch.distributed as dist

from ._lib import DfnError
from .functional import to8b


def shard_frames(n_frames, rank, world_size):
    """Contiguous frame block of `rank`: the first n % G ranks render one frame more."""
    base, extra = divmod(n_frames, world_size)
    b = rank * base + min(rank, extra)
    return b, b + base + (1 if rank < extra else 0)


class FrameSink:
    """Double-buffered device->host path for rendered frames: to8b on the render stream, copy on a side stream.
    `planes` images per frame (render_person keeps head and person).  `on_frame(k, u8[planes,H,W,3] numpy view)` is
    called on the host, in order, as soon as local frame k's copy has landed -- i.e. while later frames render."""

    def __init__(self, n_frames, H, W, device, depth=2, planes=1, on_frame=None):
        self.host = torch.empty((n_frames, planes, H, W, 3), dtype=torch.uint8).pin_memory()
        self.dev = [torch.empty((planes, H, W, 3), dtype=torch.uint8, device=device) for _ in range(depth)]
        self.copied = [None] * depth
        self.stream = torch.cuda.Stream(device=device)
        self.shape, self.i, self.delivered, self.on_frame = (planes, H, W, 3), 0, 0, on_frame

    def _deliver(self, upto):
        while self.on_frame is not None and self.delivered < upto:
            self.on_frame(self.delivered, self.host[self.delivered].numpy())
            self.delivered += 1

    def push(self, rgb_map):
        k = self.i % len(self.dev)
        cur = torch.cuda.current_stream()
        if self.copied[k] is not None:
            cur.wait_event(self.copied[k])          # the previous copy out of this buffer has finished
            if self.on_frame is not None:
                self.copied[k].synchronize()        # frame i-depth is in host memory: hand it on while frame i renders
                self._deliver(self.i - len(self.dev) + 1)
        self.dev[k].copy_(to8b(rgb_map).reshape(self.shape))
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            self.host[self.i].copy_(self.dev[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.copied[k] = ev
        self.i += 1

    def finish(self):
        self.stream.synchronize()
        self._deliver(self.i)
        out = self.host[:self.i]
        return out[:, 0] if self.shape[0] == 1 else out


@torch.no_grad()
def render_sequence(engine, H, W, focal, poses, auds, bc_rgb, near, far, cx=None, cy=None, group=None, gather=True,
                    on_frame=None):
    """FaceNeRF / NeRF sequence: poses [N,3,4] (or [N,4,4]), auds [N,dim_aud] (None for NeRF), one background.
    Returns uint8 frames [N,H,W,3] in pinned host memory (this rank's block [n_local,H,W,3] when gather=False).
    on_frame(i, u8[1,H,W,3]): called per finished frame of THIS rank with its global index (e.g. a FrameWriter)."""
    return _run(lambda i, bc, lat: engine.render_frame(H, W, focal, poses[i, :3, :4], bc, lat, near, far, cx, cy)['rgb_map'],
                H, W, poses.shape[0], bc_rgb, auds, group, gather, on_frame=on_frame)


@torch.no_grad()
def render_sequence_head_torso(decoder, H, W, focal, poses, pose_torso, bc_rgb, z_shape, z_app, signals, signals_torso,
                               near, far, cx=None, cy=None, N_samples=64, precision=None, group=None, gather=True,
                               on_frame=None, with_head=False):
    """The reference's live loop MAIN:624-733: head poses [N,3,4], one fixed body pose (MAIN:644), per-frame head
    signals [N,dim_signal] and torso signals [N,dim_et_embed].  Returns the `person` frames (MAIN:712-715) as uint8
    [N,H,W,3]; with_head=True keeps both images of a frame, [N,2,H,W,3] = (head, person) (MAIN:712-722 writes both)."""
    from .decoder import render_head_torso
    from . import _lib
    precision = _lib.PREC_BF16X3 if precision is None else precision
    lat = torch.cat([signals, signals_torso], -1)
    ds = signals.shape[1]

    def frame(i, bc, l):
        both = render_head_torso(decoder, H, W, focal, poses[i, :3, :4], pose_torso[:3, :4], bc, z_shape, z_app, l[:ds], l[ds:],
                                 near, far, cx, cy, N_samples=N_samples, precision=precision)
        return torch.cat(both, 0) if with_head else both[1]
    return _run(frame, H, W, poses.shape[0], bc_rgb, lat, group, gather, planes=2 if with_head else 1, on_frame=on_frame)


def _run(render_one, H, W, n_frames, bc_rgb, latents, group, gather, planes=1, on_frame=None):
    if not bc_rgb.is_cuda:
        raise DfnError('dfa_nerf_b200 has no CPU path: bc_rgb must be a CUDA tensor')
    device = bc_rgb.device
    multi = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if multi else 0
    world = dist.get_world_size(group) if multi else 1
    f0, f1 = shard_frames(n_frames, rank, world)
    lat_dev = latents.to(device, torch.float32).contiguous() if latents is not None else None   # one upload for the sequence
    bc = bc_rgb.reshape(-1, 3)
    squeeze = (lambda t: t[:, 0]) if planes == 1 else (lambda t: t)
    if world > 1 and gather:
        # frames stay on the device until the one all-gather at the end (uint8: 3 bytes per pixel on the wire); ragged
        # blocks are padded to the largest and trimmed afterwards
        per = (n_frames + world - 1) // world
        tile = torch.zeros((per, planes, H, W, 3), dtype=torch.uint8, device=device)
        for i in range(f0, f1):
            tile[i - f0].copy_(to8b(render_one(i, bc, lat_dev[i] if lat_dev is not None else None)).reshape(planes, H, W, 3))
        full = torch.empty((world * per, planes, H, W, 3), dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(full, tile, group=group)
        blocks = [shard_frames(n_frames, r, world) for r in range(world)]
        out = torch.cat([full[r * per:r * per + (e - b)] for r, (b, e) in enumerate(blocks)], 0).cpu()
        if on_frame is not None:
            for i in range(f0, f1):
                on_frame(i, out[i].numpy())
        return squeeze(out)
    sink = FrameSink(max(f1 - f0, 1), H, W, device, planes=planes,
                     on_frame=(lambda k, fr: on_frame(f0 + k, fr)) if on_frame is not None else None)
    for i in range(f0, f1):
        sink.push(render_one(i, bc, lat_dev[i] if lat_dev is not None else None))
    return sink.finish()[:f1 - f0]

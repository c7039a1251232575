"""Sequence-level frame loop: the caller of the hot path (MAIN:624-733).

The reference renders a driven sequence with one Python iteration per frame: latents, two ``get_rays``, a chunk
loop of ceil(HW/2048) network cascades, ``.cpu().numpy()`` + ``to8b`` + JPEG write -- each frame's device->host copy
and file I/O serialised with the next frame's compute.  Here a frame is ~10 kernel launches, so the loop is
restructured around the device:

* per-frame poses and latents are uploaded once as tables (``[N,3,4]``, ``[N,dim]``);
* the rendered frame is quantised on the device (``dfn_to8b``, HELP:17) and copied to pinned host memory as
  3 bytes/pixel on a side stream, double-buffered, so frame i's copy overlaps frame i+1's render;
* with several GPUs the FRAMES are sharded (contiguous blocks per rank); every rank's k-th frame is gathered to rank 0
  as uint8 on a communication stream while frame k+1 renders, and rank 0 alone copies the sequence out from pinned
  double buffers (FrameGather; BASELINE.json configs[4]: 300-frame sequence on 8 GPUs).

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
    `planes` images per frame (render_person keeps head and person).  `on_frame(k, u8[planes,H,W,3] numpy array)` is
    called on the host, in order, as soon as local frame k's copy has landed -- i.e. while later frames render.
    keep=True (default): the whole sequence stays in one pinned buffer, finish() returns it and on_frame gets views of it.
    keep=False: a ring of `depth + 2` pinned slots is recycled and on_frame receives its own copy of each frame (a job
    that streams thousands of frames to disk pins a few MB, not the sequence); finish() returns None."""

    def __init__(self, n_frames, H, W, device, depth=2, planes=1, on_frame=None, keep=True, host=None):
        if not keep and on_frame is None:
            raise DfnError('FrameSink(keep=False) needs an on_frame consumer')
        self.keep = keep
        self.slots = n_frames if keep else depth + 2
        if host is not None:       # caller-owned pinned sequence buffer (pinning ~180 MB costs ~0.1 s: a loop that renders many
            if not keep or not host.is_pinned() or host.dtype != torch.uint8 or host.numel() < max(self.slots, 1) * planes * H * W * 3:
                raise DfnError('FrameSink: host must be a pinned uint8 buffer of at least n_frames*planes*H*W*3 bytes (keep=True)')
            self.host = host.reshape(-1)[:max(self.slots, 1) * planes * H * W * 3].reshape(max(self.slots, 1), planes, H, W, 3)
        else:                      # sequences allocates it once and passes it in)
            self.host = torch.empty((max(self.slots, 1), planes, H, W, 3), dtype=torch.uint8).pin_memory()
        self.dev = [torch.empty((planes, H, W, 3), dtype=torch.uint8, device=device) for _ in range(depth)]
        self.copied = [None] * depth
        self.stream = torch.cuda.Stream(device=device)
        self.shape, self.i, self.delivered, self.on_frame = (planes, H, W, 3), 0, 0, on_frame

    def _deliver(self, upto):
        while self.on_frame is not None and self.delivered < upto:
            fr = self.host[self.delivered % self.slots].numpy()
            self.on_frame(self.delivered, fr if self.keep else fr.copy())
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
            # ring mode: slot (i % slots) was delivered before frame i - depth - 1 finished copying (slots = depth + 2)
            self.host[self.i % self.slots].copy_(self.dev[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.copied[k] = ev
        self.i += 1

    def finish(self):
        self.stream.synchronize()
        self._deliver(self.i)
        if not self.keep:
            return None
        out = self.host[:self.i]
        return out[:, 0] if self.shape[0] == 1 else out


class FrameGather:
    """Several ranks, frames sharded in contiguous blocks: local frame k of EVERY rank is gathered to rank `dst` as uint8
    (3 bytes per pixel on the wire) on a communication stream while frame k+1 renders, and rank `dst` copies the world's
    frames of that step out to pinned host memory from double buffers -- no rank ever holds the sequence on its device and
    only `dst` receives it.  Works on CPU tensors with gloo as well (the host-side logic test).  on_frame(i, u8[planes,H,W,3])
    is called on `dst` with the GLOBAL frame index as the steps land."""

    def __init__(self, n_frames, H, W, device, planes=1, group=None, dst=0, on_frame=None):
        self.group, self.dst_rank, self.on_frame = group, dst, on_frame
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.gdst = dist.get_global_rank(group, dst) if group is not None else dst
        self.blocks = [shard_frames(n_frames, r, self.world) for r in range(self.world)]
        self.per = max(e - b for b, e in self.blocks) if n_frames else 0
        self.cuda = torch.device(device).type == 'cuda'
        self.shape = (planes, H, W, 3)
        self.send = [torch.zeros(self.shape, dtype=torch.uint8, device=device) for _ in range(2)]
        self.is_dst = self.rank == dst
        if self.is_dst:
            self.recv = [[torch.empty(self.shape, dtype=torch.uint8, device=device) for _ in range(self.world)] for _ in range(2)]
            self.host = torch.empty((max(n_frames, 1),) + self.shape, dtype=torch.uint8)
            if self.cuda:
                self.host = self.host.pin_memory()
        self.stream = torch.cuda.Stream(device=device) if self.cuda else None
        self.done = [None, None]
        self.k, self.delivered_steps, self.n_frames = 0, 0, n_frames

    def _deliver(self, upto_step):
        while self.delivered_steps < upto_step:
            k = self.delivered_steps
            if self.is_dst and self.on_frame is not None:
                for b, e in self.blocks:
                    if b + k < e:
                        self.on_frame(b + k, self.host[b + k].numpy())
            self.delivered_steps += 1

    def push(self, u8):
        """u8: this rank's local frame k as uint8 [planes,H,W,3], or None when its block is shorter than the longest."""
        s = self.k & 1
        if self.cuda:
            cur = torch.cuda.current_stream()
            if self.done[s] is not None:
                cur.wait_event(self.done[s])         # the gather (and copy-out) that used these buffers two steps ago
                if self.is_dst and self.on_frame is not None:
                    self.done[s].synchronize()
                    self._deliver(self.k - 1)
        if u8 is not None:
            self.send[s].copy_(u8.reshape(self.shape))
        ctx = torch.cuda.stream(self.stream) if self.cuda else _Null()
        if self.cuda:
            ready = torch.cuda.Event()
            ready.record(cur)
        with ctx:
            if self.cuda:
                self.stream.wait_event(ready)
            dist.gather(self.send[s], self.recv[s] if self.is_dst else None, dst=self.gdst, group=self.group)
            if self.is_dst:
                for r, (b, e) in enumerate(self.blocks):
                    if b + self.k < e:
                        self.host[b + self.k].copy_(self.recv[s][r], non_blocking=True)
            if self.cuda:
                ev = torch.cuda.Event()
                ev.record(self.stream)
                self.done[s] = ev
        self.k += 1

    def finish(self):
        if self.cuda:
            self.stream.synchronize()
        self._deliver(self.k)
        return self.host[:self.n_frames] if self.is_dst else None


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


@torch.no_grad()
def render_sequence(engine, H, W, focal, poses, auds, bc_rgb, near, far, cx=None, cy=None, group=None, gather=True,
                    on_frame=None, keep=True, shard='frames', out=None):
    """FaceNeRF / NeRF sequence: poses [N,3,4] (or [N,4,4]), auds [N,dim_aud] (None for NeRF), one background.
    One process: returns uint8 frames [N,H,W,3] in pinned host memory.  Several ranks (torch.distributed initialised): the
    frames are sharded in contiguous blocks; gather=True delivers the whole sequence to rank 0 (return value there, None on
    the other ranks), gather=False returns this rank's block [n_local,H,W,3].
    on_frame(i, u8[1,H,W,3]): called per finished frame with its global index (e.g. a FrameWriter) -- on this rank for its
    own frames (gather=False) or on rank 0 for all of them (gather=True).  keep=False (needs on_frame): stream through a
    small pinned ring instead of keeping the sequence, returns None.
    shard='rays' (several ranks): every FRAME is split over the ranks by rays instead -- each rank renders 1/G of every frame, the
    uint8 tiles are all-gathered and rank 0 copies the frame out on a side stream while the next frame renders
    (distributed.RayShardSink).  Same frames, 1/G of the per-frame latency (a live loop), and no tail imbalance when G does not
    divide the frame count; rank 0 returns the sequence, the other ranks None.
    out: a caller-owned PINNED uint8 buffer of at least N*H*W*3 bytes that receives the frames (one process, or shard='rays' on rank 0)
    instead of a freshly pinned one."""
    poses_host = torch.as_tensor(poses, dtype=torch.float32).detach().cpu()     # ONE device->host copy for the sequence: get_rays
    #                                                                             takes the pose as kernel arguments
    return _run(lambda i, bc, lat, rr=None: engine.render_frame(H, W, focal, poses_host[i, :3, :4], bc, lat, near, far, cx, cy,
                                                                ray_range=rr)['rgb_map'],
                H, W, poses.shape[0], bc_rgb, auds, group, gather, on_frame=on_frame, keep=keep, shard=shard, out=out)


@torch.no_grad()
def render_sequence_head_torso(decoder, H, W, focal, poses, pose_torso, bc_rgb, z_shape, z_app, signals, signals_torso,
                               near, far, cx=None, cy=None, N_samples=64, precision=None, group=None, gather=True,
                               on_frame=None, with_head=False, keep=True, shard='frames'):
    """The reference's live loop MAIN:624-733: head poses [N,3,4], one fixed body pose (MAIN:644), per-frame head
    signals [N,dim_signal] and torso signals [N,dim_et_embed].  Returns the `person` frames (MAIN:712-715) as uint8
    [N,H,W,3]; with_head=True keeps both images of a frame, [N,2,H,W,3] = (head, person) (MAIN:712-722 writes both)."""
    from .decoder import render_head_torso
    from . import _lib
    precision = _lib.PREC_BF16X3 if precision is None else precision
    from .functional import get_rays
    lat = torch.cat([signals, signals_torso], -1)
    ds = signals.shape[1]
    poses_host = torch.as_tensor(poses, dtype=torch.float32).detach().cpu()     # one device->host copy, not two syncs per frame
    # the body pose is fixed (MAIN:644): its rays are computed once for the sequence
    rays_torso = [t.reshape(-1, 3) for t in get_rays(H, W, focal, torch.as_tensor(pose_torso).detach().cpu()[:3, :4], cx, cy,
                                                      device=bc_rgb.device)]

    def frame(i, bc, l, rr=None):
        both = render_head_torso(decoder, H, W, focal, poses_host[i, :3, :4], None, bc, z_shape, z_app, l[:ds], l[ds:],
                                 near, far, cx, cy, N_samples=N_samples, precision=precision, rays_torso=rays_torso, ray_range=rr)
        return torch.cat(both, 0) if with_head else both[1]
    if shard == 'rays' and with_head:
        raise DfnError('render_sequence_head_torso: shard="rays" delivers the person image only (with_head=False)')
    return _run(frame, H, W, poses.shape[0], bc_rgb, lat, group, gather, planes=2 if with_head else 1, on_frame=on_frame, keep=keep,
                shard=shard)


def _run_ray_sharded(render_one, H, W, n_frames, bc, lat_dev, device, group, rank, world, on_frame, keep, host=None):
    """shard='rays': every rank renders its ray range of EVERY frame; RayShardSink gathers the uint8 tiles and copies the frame out
    on rank 0 while the next frames render; the host hands frame i-2 on (on_frame) while frames i-1 and i are in flight."""
    from .distributed import RayShardSink, shard_range
    n = H * W
    b, e, _ = shard_range(n, rank, world)
    sink = RayShardSink(n, device, channels=3, dtype=torch.uint8, group=group, to_host=True, depth=3)   # the host runs two frames ahead:
    #                                  the ranks are coupled through the gather, one frame of slack left them waiting for the slowest host
    out = None
    if rank == 0 and keep:
        if host is not None:
            if not host.is_pinned() or host.dtype != torch.uint8 or host.numel() < max(n_frames, 1) * H * W * 3:
                raise DfnError('render_sequence: out must be a pinned uint8 buffer of at least N*H*W*3 bytes')
            out = host.reshape(-1)[:max(n_frames, 1) * H * W * 3].reshape(max(n_frames, 1), H, W, 3)
        else:
            out = torch.empty((max(n_frames, 1), H, W, 3), dtype=torch.uint8).pin_memory()

    def deliver(i):
        h = sink.wait(i)
        if rank == 0 and on_frame is not None:
            fr = h.reshape(1, H, W, 3).numpy()
            on_frame(i, fr if keep else fr.copy())

    for i in range(n_frames):
        tile = render_one(i, bc, lat_dev[i] if lat_dev is not None else None, (b, e)) if e > b else \
            torch.zeros((0, 3), dtype=torch.float32, device=device)
        sink.push(to8b(tile), host_out=out[i] if out is not None else None)
        if i > 1:
            deliver(i - 2)
    for i in range(max(n_frames - 2, 0), n_frames):
        deliver(i)
    sink.finish()
    return out[:n_frames] if out is not None else None


def _run(render_one, H, W, n_frames, bc_rgb, latents, group, gather, planes=1, on_frame=None, keep=True, shard='frames', out=None):
    if not bc_rgb.is_cuda:
        raise DfnError('dfa_nerf_b200 has no CPU path: bc_rgb must be a CUDA tensor')
    if shard not in ('frames', 'rays'):
        raise DfnError('shard must be "frames" or "rays"')
    device = bc_rgb.device
    multi = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if multi else 0
    world = dist.get_world_size(group) if multi else 1
    f0, f1 = shard_frames(n_frames, rank, world)
    lat_dev = latents.to(device, torch.float32).contiguous() if latents is not None else None   # one upload for the sequence
    bc = bc_rgb.reshape(-1, 3)
    if shard == 'rays' and world > 1:
        if not keep and on_frame is None:
            raise DfnError('keep=False needs an on_frame consumer')
        return _run_ray_sharded(render_one, H, W, n_frames, bc, lat_dev, device, group, rank, world, on_frame, keep, host=out)
    squeeze = (lambda t: t[:, 0]) if planes == 1 else (lambda t: t)
    if world > 1 and gather:
        fg = FrameGather(n_frames, H, W, device, planes=planes, group=group, on_frame=on_frame)
        for k in range(fg.per):
            i = f0 + k
            fg.push(to8b(render_one(i, bc, lat_dev[i] if lat_dev is not None else None)) if i < f1 else None)
        out = fg.finish()
        return squeeze(out) if out is not None else None
    sink = FrameSink(max(f1 - f0, 1), H, W, device, planes=planes, keep=keep, host=out if world == 1 else None,
                     on_frame=(lambda k, fr: on_frame(f0 + k, fr)) if on_frame is not None else None)
    for i in range(f0, f1):
        sink.push(render_one(i, bc, lat_dev[i] if lat_dev is not None else None))
    out = sink.finish()
    return out[:f1 - f0] if out is not None else None

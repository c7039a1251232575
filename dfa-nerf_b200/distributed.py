"""Ray sharding across the GPUs of one box.

Every ray is independent, so rank g renders the contiguous ray range
[g*ceil(HW/G), (g+1)*ceil(HW/G)) with replicated weights, and the only data-path collective is
ONE all-gather of the [ceil(HW/G), 3] fp32 RGB tile per frame (NCCL over NVLink; gloo in the CPU
tests).  The reference itself is single-GPU (scripts/test_obama.sh:1); this is new.

`gather_rgb` is the blocking form (the frame is on every rank when the call returns to the stream).  `RayShardSink`
is the pipelined form a frame LOOP uses: frame i's all-gather and rank 0's device->host copy run on a communication
stream out of double buffers while frame i+1 renders, so a ray-sharded sequence delivers frames at the kernels' rate
with one frame of latency (G GPUs cut a frame's latency by G; frame sharding, sequence.py, only raises throughput).
"""
import torch
import torch.distributed as dist


def shard_range(n_rays, rank, world_size):
    """Contiguous range of rank `rank`; all ranks get ceil(n/G) rays except the tail, which may be short/empty."""
    per = (n_rays + world_size - 1) // world_size
    b = min(rank * per, n_rays)
    e = min(b + per, n_rays)
    return b, e, per


def gather_rgb(local_rgb, n_rays, group=None):
    """all_gather_into_tensor of the padded per-rank tile -> [n_rays, C] on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_rgb[:n_rays]
    per = (n_rays + world - 1) // world
    c = local_rgb.shape[1]
    tile = local_rgb
    if tile.shape[0] != per:
        tile = torch.zeros((per, c), dtype=local_rgb.dtype, device=local_rgb.device)
        tile[:local_rgb.shape[0]] = local_rgb
    full = torch.empty((world * per, c), dtype=local_rgb.dtype, device=local_rgb.device)
    dist.all_gather_into_tensor(full, tile.contiguous(), group=group)
    return full[:n_rays]


def render_sharded(engine, H, W, focal, c2w, bc_rgb, aud, near, far, cx=None, cy=None, group=None, gather=True):
    """Renders this rank's ray range of the frame and (optionally) all-gathers the RGB image [H*W,3]."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n = H * W
    b, e, _ = shard_range(n, rank, world)
    if e > b:
        local = engine.render_frame(H, W, focal, c2w, bc_rgb, aud, near, far, cx, cy, ray_range=(b, e))['rgb_map']
    else:
        local = torch.zeros((0, 3), dtype=torch.float32, device=bc_rgb.device)
    if not gather:
        return local
    return gather_rgb(local, n, group)


class RayShardSink:
    """Output side of a ray-sharded frame loop.  push(local_rgb) takes this rank's [rows <= ceil(n/G), C] tile of frame i on
    the current stream and returns i; the all-gather of the tiles (world > 1) and -- on rank `dst`, when to_host -- the copy of
    the assembled [n_rays, C] frame into pinned host memory run on a side stream, double-buffered: buffer i % depth is reused
    by frame i + depth only after frame i's gather / copy has finished (stream-side wait) and, for the host slot, after the
    caller has taken frame i (wait(i) / lag-one consumption in frame order).  wait(i) blocks the HOST until frame i has landed
    and returns it (rank dst: the pinned [n_rays, C] slot; other ranks: None); device(i) is the gathered frame on this rank's
    device.  CPU tensors + gloo run the same logic without streams (the host-side test)."""

    def __init__(self, n_rays, device, channels=3, dtype=torch.float32, group=None, dst=0, depth=2, to_host=True):
        self.multi = dist.is_available() and dist.is_initialized()
        self.group = group
        self.rank = dist.get_rank(group) if self.multi else 0
        self.world = dist.get_world_size(group) if self.multi else 1
        self.n, self.c, self.depth = n_rays, channels, depth
        self.per = (n_rays + self.world - 1) // self.world
        self.cuda = torch.device(device).type == 'cuda'
        self.is_dst = self.rank == dst
        self.to_host = to_host
        self.send = [torch.zeros((self.per, channels), dtype=dtype, device=device) for _ in range(depth)]
        self.full = [torch.empty((self.world * self.per, channels), dtype=dtype, device=device) for _ in range(depth)] \
            if self.world > 1 else self.send
        self.host = None
        if to_host and self.is_dst:
            self.host = [torch.empty((n_rays, channels), dtype=dtype) for _ in range(depth)]
            if self.cuda:
                self.host = [h.pin_memory() for h in self.host]
        self.stream = torch.cuda.Stream(device=device) if self.cuda else None
        self.done = [None] * depth
        self.ext = [None] * depth
        self.i = 0

    def push(self, local_rgb, host_out=None):
        """host_out (rank dst): a pinned [n_rays, C] tensor that receives this frame instead of the ring slot (a sequence buffer
        the caller keeps); wait(i) then only synchronises."""
        i, s = self.i, self.i % self.depth
        rows = local_rgb.shape[0]
        if rows > self.per:
            raise ValueError('RayShardSink.push: tile has %d rows, a rank holds at most %d' % (rows, self.per))
        if self.cuda:
            cur = torch.cuda.current_stream()
            if self.done[s] is not None:
                cur.wait_event(self.done[s])         # the gather / copy that read these buffers `depth` frames ago
        self.send[s][:rows].copy_(local_rgb)
        if self.cuda:
            ready = torch.cuda.Event()
            ready.record(cur)
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ready)
                self._finish_frame(s, host_out)
                ev = torch.cuda.Event()
                ev.record(self.stream)
            self.done[s] = ev
        else:
            self._finish_frame(s, host_out)
        self.ext[s] = host_out if self.is_dst else None
        self.i += 1
        return i

    def _finish_frame(self, s, host_out=None):
        if self.world > 1:
            dist.all_gather_into_tensor(self.full[s], self.send[s], group=self.group)
        if self.is_dst and host_out is not None:
            host_out.copy_(self.full[s][:self.n].reshape(host_out.shape), non_blocking=True)
        elif self.host is not None:
            self.host[s].copy_(self.full[s][:self.n], non_blocking=True)

    def _check(self, i):
        if not (self.i - self.depth <= i < self.i) or i < 0:
            raise IndexError('RayShardSink: frame %d is not in flight (frames %d..%d are)' % (i, max(self.i - self.depth, 0), self.i - 1))
        return i % self.depth

    def wait(self, i):
        s = self._check(i)
        if self.cuda and self.done[s] is not None:
            self.done[s].synchronize()
        if self.ext[s] is not None:
            return self.ext[s]
        return self.host[s] if self.host is not None else None

    def device(self, i):
        s = self._check(i)
        if self.cuda and self.done[s] is not None:
            torch.cuda.current_stream().wait_event(self.done[s])
        return self.full[s][:self.n]

    def finish(self):
        if self.cuda:
            self.stream.synchronize()

"""Ray sharding across the GPUs of one box.

Every ray is independent, so rank g renders the contiguous ray range
[g*ceil(HW/G), (g+1)*ceil(HW/G)) with replicated weights, and the only data-path collective is
ONE all-gather of the [ceil(HW/G), 3] fp32 RGB tile per frame (NCCL over NVLink; gloo in the CPU
tests).  The reference itself is single-GPU (scripts/test_obama.sh:1); this is new.
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

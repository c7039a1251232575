"""Under torchrun (N >= 2): the ray-sharded sequence loop (render_sequence(..., shard='rays')) against the frame-sharded one and,
on rank 0, against per-frame renders -- uint8 frames must be equal -- and the time of both.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 profiles/check_ray_sharded_sequence.py [frames]"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfa_nerf_b200 as dfn  # noqa: E402
import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dev = torch.device('cuda', lr)
dist.init_process_group('nccl', device_id=dev)
H = W = 450


def mk(seed):
    m = dfn.FaceNeRF(D=8, W=256, input_ch=63, input_ch_views=27, dim_aud=64, output_ch=4, skips=[4], use_viewdirs=True)
    m.load_state_dict(synth.facenerf_state_dict(seed))
    return m.to(dev)


eng = dfn.RenderEngine(mk(0), mk(1), 64, 128, precision=dfn.PREC_BF16)
seq = synth.frame_inputs(H=H, W=W, seed=0, n_frames=n)
bc = seq['bc_rgb'].to(dev)
out = {}
for shard in ('frames', 'rays', 'frames', 'rays'):
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    out[shard] = dfn.render_sequence(eng, H, W, seq['focal'], seq['c2w_seq'], seq['aud'], bc, seq['near'], seq['far'], seq['cx'], seq['cy'],
                                     shard=shard)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print('shard=%-6s %d frames on %d GPUs: %.1f ms (%.2f ms/frame)' % (shard, n, world, (time.perf_counter() - t0) * 1e3,
                                                                           (time.perf_counter() - t0) * 1e3 / n), flush=True)
if rank == 0:
    a, b = out['frames'], out['rays']
    print('ray-sharded == frame-sharded (uint8, %s): %s' % (tuple(b.shape), bool(torch.equal(a, b))))
    ref = dfn.to8b(eng.render_frame(H, W, seq['focal'], seq['c2w_seq'][n - 1, :3, :4], bc, seq['aud'][n - 1].to(dev), seq['near'], seq['far'],
                                    seq['cx'], seq['cy'])['rgb_map']).reshape(H, W, 3).cpu()
    print('last frame == single-GPU render of that frame: %s' % bool(torch.equal(b[n - 1], ref)))
else:
    assert out['rays'] is None
dist.destroy_process_group()

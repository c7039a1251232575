"""CPU, world_size 2, gloo: the host-side sharding logic of the N>1 path (shard_range + one all-gather of the
RGB tile).  The per-rank render is the oracle here (the product has no CPU path); on GPUs the same functions run
with NCCL (bench.py --gpus N)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, W, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from dfa_nerf_b200.distributed import shard_range, gather_rgb
        from oracle import nerf_oracle as O, synth
        torch.set_num_threads(2)
        fr = synth.frame_inputs(H=H, W=W, seed=4)
        sd_c, sd_f = synth.facenerf_state_dict(0), synth.facenerf_state_dict(1)
        n = H * W
        b, e, per = shard_range(n, rank, world)
        with torch.no_grad():
            local = O.render(H, W, fr['focal'], fr['cx'], fr['cy'], fr['c2w'], fr['bc_rgb'], fr['aud'], sd_c, sd_f,
                             fr['near'], fr['far'], 16, 16, chunk=64, ray_slice=(b, e))['rgb_map']
        full = gather_rgb(local, n)
        if rank == 0:
            with torch.no_grad():
                ref = O.render(H, W, fr['focal'], fr['cx'], fr['cy'], fr['c2w'], fr['bc_rgb'], fr['aud'], sd_c, sd_f,
                               fr['near'], fr['far'], 16, 16, chunk=64)['rgb_map']
            q.put((tuple(full.shape), float((full - ref).abs().max()), per, e - b))
    finally:
        dist.destroy_process_group()


def test_ray_sharding_two_ranks_gloo():
    H, W = 9, 7            # 63 rays: ragged split 32 + 31, padded tile in the gather
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, H, W, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    shape, err, per, mine = q.get(timeout=10)
    assert shape == (H * W, 3) and per == 32 and mine == 32
    assert err == 0.0      # rays are independent: sharded == unsharded bit for bit


def test_gather_rgb_single_process():
    from dfa_nerf_b200.distributed import gather_rgb
    x = torch.rand(10, 3)
    assert torch.equal(gather_rgb(x, 10), x)


def test_shard_frames_partition():
    """Frame sharding of the sequence loop: contiguous, disjoint, complete, balanced to within one frame."""
    from dfa_nerf_b200.sequence import shard_frames
    for n in (0, 1, 7, 300):
        for world in (1, 2, 3, 8):
            blocks = [shard_frames(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            sizes = [e - b for b, e in blocks]
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n
    assert shard_frames(300, 7, 8) == (263, 300) and shard_frames(300, 0, 8) == (0, 38)


def _frame(i, H=12, W=10):
    import numpy as np
    y, x = np.mgrid[0:H, 0:W]
    return np.stack([(x * 20 + i) % 256, (y * 15 + 2 * i) % 256, np.full_like(x, (40 * i) % 256)], -1).astype(np.uint8)


def _writer_worker(rank, world, port, n_frames, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from dfa_nerf_b200.sequence import shard_frames
        from dfa_nerf_b200.frame_io import FrameWriter
        b, e = shard_frames(n_frames, dist.get_rank(), dist.get_world_size())
        w = FrameWriter(os.path.join(out_dir, 'render_com'), workers=2)
        for i in range(b, e):                 # what sequence._run's on_frame hands over: the GLOBAL frame index
            w.write(i, _frame(i))
        paths = w.close()
        assert len(paths) == e - b
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_frame_sharded_ranks_fill_one_directory(tmp_path):
    """Output side of the frame-sharded sequence (render_person with several ranks): every rank writes its own block of
    frames under their global numbers -- one directory, no collisions, no gather."""
    import numpy as np
    from PIL import Image
    n = 7                                      # ragged: 4 + 3
    ctx = mp.get_context('spawn')
    port = _free_port()
    procs = [ctx.Process(target=_writer_worker, args=(r, 2, port, n, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    names = sorted(os.listdir(tmp_path / 'render_com'))
    assert names == ['test_%06d.jpg' % i for i in range(n)]
    for i, name in enumerate(names):
        back = np.asarray(Image.open(tmp_path / 'render_com' / name)).astype(np.int32)
        assert back.shape == (12, 10, 3) and np.abs(back - _frame(i)).mean() < 12.0      # the right frame under each number


def _gather_worker(rank, world, port, n_frames, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from dfa_nerf_b200.sequence import shard_frames, FrameGather
        seen = []
        fg = FrameGather(n_frames, 12, 10, 'cpu', planes=1, on_frame=lambda i, fr: seen.append((i, int(fr.sum()))))
        b, e = shard_frames(n_frames, rank, world)
        assert fg.per == -(-n_frames // world)
        for k in range(fg.per):                     # every rank makes the same number of collective calls
            i = b + k
            fg.push(torch.from_numpy(_frame(i))[None] if i < e else None)
        out = fg.finish()
        if rank == 0:
            q.put((tuple(out.shape), [int(out[i].sum()) for i in range(n_frames)], sorted(seen)))
        else:
            assert out is None and seen == []
    finally:
        dist.destroy_process_group()


def test_frame_gather_to_rank0_two_ranks_gloo():
    """The world>1 sequence path (sequence.FrameGather): frames sharded in contiguous blocks, local frame k of every rank
    gathered to rank 0 as uint8 per step, ragged blocks (4 + 3) padded with a step whose stale buffer is ignored."""
    n = 7
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    shape, sums, seen = q.get(timeout=10)
    want = [int(_frame(i).sum()) for i in range(n)]
    assert shape == (n, 1, 12, 10, 3) and sums == want
    assert seen == [(i, want[i]) for i in range(n)]          # on_frame saw every global frame exactly once, on rank 0


def _sink_worker(rank, world, port, n, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from dfa_nerf_b200.distributed import RayShardSink, shard_range
        b, e, per = shard_range(n, rank, world)
        sink = RayShardSink(n, 'cpu', depth=2)
        frames = [torch.arange(n * 3, dtype=torch.float32).reshape(n, 3) * (k + 1) for k in range(5)]
        ok = True
        for k, f in enumerate(frames):
            i = sink.push(f[b:e])
            ok &= i == k
            if i > 0:                                   # lag-one consumption, as the frame loop does
                h = sink.wait(i - 1)
                ok &= (h is not None) == (rank == 0)
                if h is not None:
                    ok &= torch.equal(h, frames[i - 1])
            ok &= torch.equal(sink.device(i), f)        # every rank holds the gathered frame
        sink.finish()
        # uint8 tiles and a caller-owned sequence buffer (the ray-sharded sequence loop, sequence._run_ray_sharded)
        seq = torch.zeros((3, 7, 9, 3), dtype=torch.uint8) if rank == 0 else None
        s8 = RayShardSink(n, 'cpu', dtype=torch.uint8)
        f8 = [(torch.arange(n * 3, dtype=torch.int64).reshape(n, 3) * (k + 3) % 251).to(torch.uint8) for k in range(3)]
        for k, f in enumerate(f8):
            s8.push(f[b:e], host_out=seq[k] if seq is not None else None)
            got = s8.wait(k)
            ok &= (got is not None) == (rank == 0)
            if got is not None:
                ok &= got.data_ptr() == seq[k].data_ptr() and torch.equal(seq[k].reshape(n, 3), f)
        s8.finish()
        try:
            sink.wait(1)                                # long recycled
            ok = False
        except IndexError:
            pass
        q.put((rank, bool(ok), per))
    finally:
        dist.destroy_process_group()


def test_ray_shard_sink_two_ranks_gloo():
    """RayShardSink (the pipelined gather + copy-out of a ray-sharded frame loop): frame order, ragged tiles, double-buffer
    reuse, rank 0 alone receives host frames."""
    n = 63                                             # 32 + 31
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sink_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=10) for _ in range(2))
    assert got == [(0, True, 32), (1, True, 32)]


def test_ray_shard_sink_single_process():
    from dfa_nerf_b200.distributed import RayShardSink
    sink = RayShardSink(10, 'cpu')
    x = torch.rand(10, 3)
    i = sink.push(x)
    assert torch.equal(sink.wait(i), x) and torch.equal(sink.device(i), x)
    import pytest
    with pytest.raises(ValueError):
        sink.push(torch.rand(11, 3))

#!/usr/bin/env python
"""bench.py -- rendered rays/s of the hierarchical NeRF hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16|bf16x3|fp32]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference            # CPU arm: the oracle port on the host cores
    python bench.py --workload head_torso       # BASELINE.json configs[2]: the reference's live two-field frame
    python bench.py --workload mlp_1m           # configs[3]: flat 8x256 NeRF query, 2^20 rays x 192 samples (roofline sweep)
    python bench.py --workload sequence --steps 1   # configs[4]: driven sequence (--frames N, default 300), frames sharded

One step = one 450x450 frame (202,500 rays) x (64 coarse + 128 fine samples) of the synthetic
FaceNeRF field (configs[1] of BASELINE.json): get_rays -> z sampling -> PE + 8x256 skip-MLP ->
compositing -> sample_pdf -> sort-merge -> fine MLP -> RGB.  With N > 1 the frame's rays are sharded
contiguously over the ranks and the RGB tile is all-gathered (NCCL) inside the timed step.

Printed JSON (rank 0): `value` = rays/s with inputs resident in HBM; `e2e` = the same through the
public API with pinned-host inputs/outputs copied every step; `roofline` = tcgen05 MLP kernel against
the measured bf16 peak; `cpu_baseline` = the oracle port on this box's host cores (bounded sample).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'rendered rays/sec at 450x450x(64+128) samples; 1/2/4/8 B200 vs CPU ref'
H = W = 450
N_SAMPLES, N_IMPORTANCE = 64, 128


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_burst=d['bf16_tflops'], bf16_sustained=d.get('bf16_tflops_sustained', d['bf16_tflops']),
                    hbm=d['hbm_gbs'], source='measured (MEASURED_PEAKS.json)')
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source='fallback (B200_PROFILING.md)')


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t0 = index, [], None, 0.0

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '25'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(',')]))

    def mark(self):
        """Samples before this point (sampler start-up, idle GPU) are not part of the record."""
        self.t0 = time.time()

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        rows = [r for t, r in self.rows if t >= self.t0] or [r for _, r in self.rows]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def cpu_arm_head_torso(steps, warmup, rays_per_step, threads=None):
    """The reference's live chunk (MAIN:633-708: Decoder head + torso with deformation, two-field compositing) on the
    host cores: oracle port, fp32, 64 samples, chunk=2048."""
    import torch
    from oracle import nerf_oracle as O, synth
    torch.set_num_threads(threads or os.cpu_count() or 1)
    fr, fr_t = synth.frame_inputs(H=H, W=W, seed=0), synth.frame_inputs(H=H, W=W, seed=7)
    sd = synth.decoder_state_dict(0)
    g = torch.Generator().manual_seed(0)
    zs, za = torch.randn(1, 2, 256, generator=g), torch.randn(1, 2, 256, generator=g)
    sig, sig_t = torch.randn(1, 96, generator=g), torch.randn(1, 42, generator=g)
    b = (H * W) // 2 - rays_per_step // 2
    ro, rd = [t.reshape(-1, 3) for t in O.get_rays(H, W, fr['focal'], fr['c2w'], fr['cx'], fr['cy'])]
    rot, rdt = [t.reshape(-1, 3) for t in O.get_rays(H, W, fr['focal'], fr_t['c2w'], fr['cx'], fr['cy'])]
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            for c0 in range(b, b + rays_per_step, 2048):
                c1 = min(c0 + 2048, b + rays_per_step)
                z = O.z_vals_uniform(torch.full((c1 - c0, 1), fr['near']), torch.full((c1 - c0, 1), fr['far']), N_SAMPLES)
                O.render_head_torso_chunk(sd, ro[c0:c1], rd[c0:c1], rot[c0:c1], rdt[c0:c1], z, fr['bc_rgb'][c0:c1], zs, za, sig, sig_t)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return rays_per_step / med, med, torch.get_num_threads()


def cpu_arm(steps, warmup, rays_per_step, threads=None):
    """The reference's arithmetic (oracle port: HELP.get_rays/Embedder/FaceNeRF/sample_pdf + MAIN.calc_volume_weights
    composed in upstream render() order) on the host cores, fp32, chunk=2048, bounded ray sample per step."""
    import torch
    from oracle import nerf_oracle as O, synth
    torch.set_num_threads(threads or os.cpu_count() or 1)
    fr = synth.frame_inputs(H=H, W=W, seed=0)
    sd_c, sd_f = synth.facenerf_state_dict(0), synth.facenerf_state_dict(1)
    b = (H * W) // 2 - rays_per_step // 2          # centre of the image: foreground rays
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.render(H, W, fr['focal'], fr['cx'], fr['cy'], fr['c2w'], fr['bc_rgb'], fr['aud'], sd_c, sd_f,
                     fr['near'], fr['far'], N_SAMPLES, N_IMPORTANCE, chunk=2048, ray_slice=(b, b + rays_per_step))
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return rays_per_step / med, med, torch.get_num_threads()


WORKLOADS = {
    'facenerf': 'FaceNeRF 450x450 x (64 coarse + 128 fine samples), synthetic seeded weights/pose/latent '
                '(BASELINE.json configs[1]); one step = one frame',
    'head_torso': 'Decoder head + torso (DeformationField_ori) two-field frame, 450x450 x 64 samples, synthetic seeded '
                  'weights/poses/latents (BASELINE.json configs[2], the path scripts/test_obama.sh runs); one step = one frame',
}
WORKLOADS['mlp_1m'] = 'Synthetic random-weight 8x256 NeRF (no latent), 2^20 rays x 192 samples, flat network query + compositing ' \
                      '(BASELINE.json configs[3]); one step = all rays'
WORKLOADS['sequence'] = 'Audio-driven FaceNeRF sequence, 450x450 x (64+128) per frame, per-frame pose + latent tables, uint8 frames ' \
                        'copied out double-buffered, FRAMES sharded over the GPUs (BASELINE.json configs[4]); one step = the sequence'
CPU_RAYS = int(os.environ.get('DFN_BENCH_CPU_RAYS', 8192))   # bounded CPU sample per step (four chunks of 2048 at the
                                                              # image centre; ~4 s on 16 cores; the env override is for the tests)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    rays = CPU_RAYS
    arm = cpu_arm_head_torso if args.workload == 'head_torso' else cpu_arm
    value, sec, cores = arm(max(1, args.steps), max(1, min(args.warmup, 1)), rays)
    evals = '(64+192) FaceNeRF' if args.workload == 'facenerf' else '(64 head + 64 torso) Decoder'
    sample = '%d rays (chunks of 2048, image centre) x %s evaluations per step, median of %d' % (rays, evals, max(1, args.steps))
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'rays/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOADS[args.workload] + '; CPU arm renders a bounded ray sample of the same frame',
                   'rays_per_frame': H * W, 'rays_per_step': rays},
        'cpu_baseline': {'value': value, 'unit': 'rays/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp16', 'bf16x3', 'fp32'])
    ap.add_argument('--workload', default='facenerf', choices=sorted(WORKLOADS))
    ap.add_argument('--frames', type=int, default=300, help='sequence workload: frames per step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import dfa_nerf_b200 as dfn
    from dfa_nerf_b200.distributed import shard_range, gather_rgb
    import synth                       # seeded synthetic data (repo root; nothing under oracle/ is touched on this path)

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a GPU (there is no CPU path)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    prec = {'bf16': dfn.PREC_BF16, 'fp16': dfn.PREC_FP16, 'bf16x3': dfn.PREC_BF16X3, 'fp32': dfn.PREC_FP32}[args.precision]

    fr = synth.frame_inputs(H=H, W=W, seed=0)
    n_rays = H * W
    b, e, per = shard_range(n_rays, rank, world)
    bc_dev = fr['bc_rgb'].to(dev)
    bc_host = fr['bc_rgb'].pin_memory()
    rgb_host = torch.empty((n_rays, 3), dtype=torch.float32).pin_memory()
    launches = [0]

    if args.workload == 'facenerf':
        def mk(seed):
            m = dfn.FaceNeRF(D=8, W=256, input_ch=63, input_ch_views=27, dim_aud=64, output_ch=4, skips=[4], use_viewdirs=True)
            m.load_state_dict(synth.facenerf_state_dict(seed))
            return m.to(dev)

        eng = dfn.RenderEngine(mk(0), mk(1), N_SAMPLES, N_IMPORTANCE, precision=prec)
        aud_dev = fr['aud'].to(dev)
        lat_host = fr['aud'].pin_memory()
        evals_per_ray = N_SAMPLES + N_SAMPLES + N_IMPORTANCE
        kernel_name = 'mlp_pp_kernel<bf16x3>' if prec == dfn.PREC_BF16X3 else 'mlp_tc_kernel<%s>' % args.precision
        flops_note = 'algorithmic, latent/viewdir columns folded: 2*557,184 per MLP evaluation (BASELINE.md section 2)'

        def render(bc_full, lat):
            out = eng.render_frame(H, W, fr['focal'], fr['c2w'], bc_full, lat, fr['near'], fr['far'], fr['cx'], fr['cy'],
                                   ray_range=(b, e), want=('rgb_map',))
            launches[0] += eng.last_launches + 1            # + get_rays
            return out['rgb_map']
    elif args.workload == 'mlp_1m':
        n_rays = 1 << 20
        b, e, per = shard_range(n_rays, rank, world)
        S = N_SAMPLES + N_IMPORTANCE
        net = dfn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=4, skips=[4], use_viewdirs=True)
        net.load_state_dict(synth.nerf_state_dict(0))
        net = net.to(dev)
        eng = dfn.RenderEngine(net, None, S, 0, precision=prec)
        g = torch.Generator().manual_seed(0)
        n_loc = e - b
        ro = fr['c2w'][:3, -1].expand(n_loc, 3).contiguous().to(dev)
        rd = (torch.randn(n_loc, 3, generator=g) * 0.2 + torch.tensor([0., 0., -1.])).to(dev)
        vd = rd / torch.norm(rd, dim=-1, keepdim=True)
        z, _ = torch.sort(torch.rand(n_loc, S, generator=g) * 0.6 + 0.4, -1)
        z = z.to(dev)
        bc_dev = torch.rand(n_rays, 3, generator=g).to(dev)
        bc_host = bc_dev.cpu().pin_memory()
        rgb_host = torch.empty((n_rays, 3), dtype=torch.float32).pin_memory()
        aud_dev, lat_host = torch.zeros(1, device=dev), torch.zeros(1).pin_memory()
        evals_per_ray = S
        kernel_name = 'mlp_pp_kernel<bf16x3>' if prec == dfn.PREC_BF16X3 else 'mlp_tc_kernel<%s>' % args.precision
        flops_note = 'algorithmic, viewdir columns folded: 2*557,184 per MLP evaluation (NeRF and FaceNeRF coincide, SURVEY 8d)'

        def render(bc_full, lat):
            raw = eng.query_points(net, ro, rd, vd, z, None)
            launches[0] += eng.last_launches + 1
            return dfn.raw2outputs(raw, z, rd, bc_full[b:e])[0]
    elif args.workload == 'sequence':
        def mk(seed):
            m = dfn.FaceNeRF(D=8, W=256, input_ch=63, input_ch_views=27, dim_aud=64, output_ch=4, skips=[4], use_viewdirs=True)
            m.load_state_dict(synth.facenerf_state_dict(seed))
            return m.to(dev)
        eng = dfn.RenderEngine(mk(0), mk(1), N_SAMPLES, N_IMPORTANCE, precision=prec)
        seq = synth.frame_inputs(H=H, W=W, seed=0, n_frames=args.frames)
        evals_per_ray = N_SAMPLES + N_SAMPLES + N_IMPORTANCE
        kernel_name = 'mlp_pp_kernel<bf16x3>' if prec == dfn.PREC_BF16X3 else 'mlp_tc_kernel<%s>' % args.precision
        flops_note = 'algorithmic, latent/viewdir columns folded: 2*557,184 per MLP evaluation (BASELINE.md section 2)'
        n_rays = H * W * args.frames            # rays per step: the whole sequence
        b, e = 0, n_rays
        lat_host = seq['aud'].pin_memory()
        aud_dev = None
    else:
        if prec == dfn.PREC_FP32:
            raise SystemExit('--workload head_torso runs on the tensor-core path: --precision bf16 | bf16x3')
        dec = dfn.Decoder(z_dim=256, hidden_size=256, dim_signal=96, use_deformation_field=True)
        dec.load_state_dict(synth.decoder_state_dict(0))
        dec = dec.to(dev)
        c2w_torso = synth.frame_inputs(H=H, W=W, seed=7)['c2w']        # fixed body pose (MAIN:644)
        g = torch.Generator().manual_seed(0)
        zs, za = torch.randn(1, 2, 256, generator=g).to(dev), torch.randn(1, 2, 256, generator=g).to(dev)
        lat = torch.cat([torch.randn(96, generator=g), torch.randn(42, generator=g)])   # head signal | torso signal
        aud_dev, lat_host = lat.to(dev), lat.pin_memory()
        evals_per_ray = 2 * N_SAMPLES
        kernel_name = 'mlp_pp_kernel<%s, Decoder>' % args.precision
        flops_note = 'algorithmic, per-frame latents and per-ray view term folded: 2*556,032 (head) + 2*628,352 (torso incl. ' \
                     'deformation field) per sample (SURVEY.md section 8d, appendix A)'

        def render(bc_full, lat):
            _, rgb = dfn.render_head_torso(dec, H, W, fr['focal'], fr['c2w'], c2w_torso, bc_full, zs, za, lat[:96], lat[96:],
                                           fr['near'], fr['far'], fr['cx'], fr['cy'], N_samples=N_SAMPLES, ray_range=(b, e),
                                           precision=prec)
            launches[0] += dfn.render_head_torso.last_launches
            return rgb

    def step_sequence(n=None):
        # pose / latent tables from (pinned) host memory, uint8 frames back to pinned host memory: this IS the end-to-end path
        n = args.frames if n is None else n
        frames = dfn.render_sequence(eng, H, W, seq['focal'], seq['c2w_seq'][:n], lat_host[:n], bc_dev, seq['near'], seq['far'],
                                     seq['cx'], seq['cy'])
        launches[0] += (eng.last_launches + 2) * ((n + world - 1) // world)
        return frames

    def step_resident():
        if args.workload == 'sequence':
            return step_sequence()
        rgb = render(bc_dev, aud_dev)
        return gather_rgb(rgb, n_rays) if world > 1 else rgb

    def step_e2e():
        if args.workload == 'sequence':
            return step_sequence()
        bc = bc_host[b:e].to(dev, non_blocking=True)
        lat = lat_host.to(dev, non_blocking=True)
        bc_full = torch.empty((n_rays, 3), dtype=torch.float32, device=dev)
        bc_full[b:e] = bc
        rgb = render(bc_full, lat)
        full = gather_rgb(rgb, n_rays) if world > 1 else rgb
        if rank == 0:
            rgb_host.copy_(full, non_blocking=True)
        torch.cuda.current_stream().synchronize()       # the caller consumes the frame

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps

    # ---- device-resident throughput, with per-launch events around the tcgen05 kernel ---------------
    # clocks are sampled from the warm-up through both timed regions (the GPU is under this load throughout; a
    # multi-GPU step is a few milliseconds, shorter than nvidia-smi's start-up)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    sampler.mark()
    for _ in range(args.warmup):
        if args.workload == 'sequence':
            step_sequence(min(args.frames, 2 * world))     # warm-up on a short prefix of the sequence
        else:
            step_resident()
    torch.cuda.synchronize()
    dfn.lib.dfn_profile_enable(1 if prec != dfn.PREC_FP32 else 0)
    launches[0] = 0
    ms_step = timed(step_resident, args.steps, 0)
    k_ms, k_n, k_macs = C.c_double(), C.c_int64(), C.c_double()
    dfn.lib.dfn_profile_collect(C.byref(k_ms), C.byref(k_n), C.byref(k_macs))
    dfn.lib.dfn_profile_enable(0)
    n_launch = launches[0]
    value = n_rays / (ms_step * 1e-3)

    # ---- end to end through the public API with host buffers ------------------------------------------
    ms_e2e = ms_step if args.workload == 'sequence' else timed(step_e2e, args.steps, 2)   # the sequence step IS end to end
    clocks = sampler.stop()
    e2e = {'value': n_rays / (ms_e2e * 1e-3), 'unit': 'rays/s',
           'h2d_bytes_per_step': int((e - b) * 12 + lat_host.numel() * 4 + 48),
           'd2h_bytes_per_step': int(n_rays * 12) if rank == 0 else 0}
    if args.workload == 'sequence':
        e2e.update(h2d_bytes_per_step=int(lat_host.numel() * 4 + args.frames * 48), d2h_bytes_per_step=int(n_rays * 3))

    pk = peaks()
    roofline = None
    if k_n.value > 0:
        achieved = 2.0 * k_macs.value / (k_ms.value * 1e-3) / 1e12
        roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': pk['bf16_sustained'], 'unit': 'TFLOP/s',
                    'frac': achieved / pk['bf16_sustained'], 'traffic': None,
                    'kernel': kernel_name,
                    'launches': int(k_n.value), 'avg_launch_ms': k_ms.value / k_n.value,
                    'kernel_share_of_step': k_ms.value / (ms_step * args.steps),
                    'flops': flops_note,
                    'peak_source': 'bf16 dense sustained, ' + pk['source']}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload in ('facenerf', 'head_torso'):
        v, sec, cores = (cpu_arm if args.workload == 'facenerf' else cpu_arm_head_torso)(3, 1, CPU_RAYS)
        cpu = {'value': v, 'unit': 'rays/s', 'cores': cores, 'kind': 'port',
               'sample': '%d rays (chunks of 2048, image centre) x %d network evaluations per ray, median of 3 (%.1f s each)'
                         % (CPU_RAYS, evals_per_ray, sec)}

    if rank == 0:
        print(json.dumps({
            'metric': METRIC, 'value': value, 'unit': 'rays/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': {'bf16': 'bf16', 'fp16': 'fp16', 'bf16x3': 'bf16x3 (split bf16, fp32-parity)', 'fp32': 'f32'}[args.precision],
            'data': 'synthetic',
            'config': {'workload': WORKLOADS[args.workload],
                       'rays_per_step': n_rays, 'mlp_evals_per_ray': evals_per_ray, 'precision': args.precision,
                       'parallelism': ('frames sharded over %d GPU(s), one all-gather of the uint8 frames at the end' % world)
                       if args.workload == 'sequence' else 'rays sharded over %d GPU(s), one all-gather of the RGB tile' % world,
                       'l2': 'per-step intermediates (~1.4 GB of raw/z buffers) exceed the 126 MB L2; no explicit flush'},
            'e2e': e2e, 'gpu_launches': n_launch, 'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""bench.py -- rendered rays/s of the hierarchical NeRF hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16|fp16|bf16x3|fp32]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference            # CPU arm: the reference's own code (oracle/_ref) on the host cores
    python bench.py --workload head_torso       # BASELINE.json configs[2]: the reference's live two-field frame
    python bench.py --workload coarse64         # configs[0]: 64x64 rays x 64 coarse samples (the reference's CPU-sized case)
    python bench.py --workload mlp_1m           # configs[3]: flat 8x256 NeRF query, 2^20 rays x 192 samples (roofline sweep)
    python bench.py --workload sequence --steps 1   # configs[4]: driven sequence (--frames N, default 300); N > 1: frames ray-sharded
    python bench.py --workload train_step       # the reference's training step (2048 rays x 64 samples x 2 fields,
                                                # forward + backward + Adam), value = training rays/s

One step = one 450x450 frame (202,500 rays) x (64 coarse + 128 fine samples) of the synthetic
FaceNeRF field (configs[1] of BASELINE.json): get_rays -> z sampling -> PE + 8x256 skip-MLP ->
compositing -> sample_pdf -> sort-merge -> fine MLP -> RGB.  With N > 1 the frame's rays are sharded
contiguously over the ranks and every frame's RGB tiles are all-gathered (NCCL) inside the timed region -- on a side
stream, frame i's gather under frame i+1's kernels (dfa_nerf_b200.RayShardSink).

Printed JSON (rank 0), one line:
  value         rays/s with inputs resident in HBM             e2e      the same through the public API, pinned-host
  roofline      tcgen05 MLP kernel vs the measured bf16 peak            inputs / outputs copied every step (the frame's
                                                                        D2H triple-buffered; the host takes frame i-2)
  modes         the same frame in the other tensor-core precisions (value, e2e, roofline fraction) -- bf16x3 is the
                mode that meets the 1e-4 float tolerance, bf16 the one BASELINE.json's config names
  parity        max-abs RGB error of every precision against the CPU arm's render of the same rays: teacher-forced (the
                reference's z_samples injected) and free-running
  extra         (default workload only) the other BASELINE.json configs measured in the same process: head_torso
                (configs[2]), sequence_300 (configs[4]), coarse64 (configs[0])
  cpu_baseline  the reference's own code (kind "reference", oracle/_ref) or the oracle port on this box's host cores
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
PRECISIONS = ('bf16', 'fp16', 'bf16x3', 'fp16x3m', 'fp32')


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_burst=d['bf16_tflops'], bf16_sustained=d.get('bf16_tflops_sustained', d['bf16_tflops']),
                    hbm=d['hbm_gbs'], source='measured (MEASURED_PEAKS.json)')
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source='fallback (B200_PROFILING.md)')


def ncu_traffic():
    """Per-point DRAM and L2->SM bytes of the tcgen05 kernels from the committed ncu --set full captures
    (profiles/traffic.json, written by profiles/summarize.py); None when absent."""
    p = os.path.join(ROOT, 'profiles', 'traffic.json')
    return json.load(open(p)) if os.path.exists(p) else {}


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

    def stop(self, t1=None):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        t1 = t1 or time.time() + 1
        rows = [r for t, r in self.rows if self.t0 <= t <= t1] or [r for _, r in self.rows]
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


# --------------------------------------------------------------------------------------------- CPU arms
CPU_RAYS = int(os.environ.get('DFN_BENCH_CPU_RAYS', 8192))   # bounded CPU sample per step (four chunks of 2048 at the
                                                              # image centre; ~4 s on 16 cores; the env override is for the tests)


def _head_torso_latents():
    import torch
    g = torch.Generator().manual_seed(0)
    zs, za = torch.randn(1, 2, 256, generator=g), torch.randn(1, 2, 256, generator=g)
    return zs, za, torch.randn(1, 96, generator=g), torch.randn(1, 42, generator=g)


def cpu_arm(workload, steps, warmup, rays_per_step, threads=None):
    """The reference's arithmetic on the host cores, fp32, chunk = 2048, a bounded ray range of the workload's frame per step.
    kind "reference": the reference's own modules vendored under oracle/_ref (oracle/ref_arm.py); kind "port": the
    restatement in oracle/nerf_oracle.py (bit-equal by oracle/make_golden.py) when oracle/_ref is absent.
    Returns dict(value rays/s, sec median, cores, kind, slice (b, e), out {rgb_map[, rgb0, z_samples]})."""
    import torch
    from oracle import nerf_oracle as O, ref_arm
    import synth
    torch.set_num_threads(threads or os.cpu_count() or 1)
    use_ref = ref_arm.available() and os.environ.get('DFN_BENCH_CPU_KIND', 'reference') != 'port'
    hw = 64 if workload == 'coarse64' else H
    fr = synth.frame_inputs(H=hw, W=hw, seed=0)
    n = hw * hw
    rays_per_step = min(rays_per_step, n)
    b = n // 2 - rays_per_step // 2          # centre of the image: foreground rays
    e = b + rays_per_step
    if workload == 'head_torso':
        sd = synth.decoder_state_dict(0)
        zs, za, sig, sig_t = _head_torso_latents()
        c2w_t = synth.frame_inputs(H=hw, W=hw, seed=7)['c2w']
        if use_ref:
            model = ref_arm.HeadTorsoFrame(sd)

            def run():
                h, p = model.render(hw, hw, fr['focal'], fr['cx'], fr['cy'], fr['c2w'], c2w_t, fr['bc_rgb'], zs, za, sig, sig_t,
                                    fr['near'], fr['far'], (b, e), N_SAMPLES)
                return {'rgb_head': h, 'rgb_map': p}
        else:
            ro, rd = [t.reshape(-1, 3) for t in O.get_rays(hw, hw, fr['focal'], fr['c2w'], fr['cx'], fr['cy'])]
            rot, rdt = [t.reshape(-1, 3) for t in O.get_rays(hw, hw, fr['focal'], c2w_t, fr['cx'], fr['cy'])]

            def run():
                hs, ps = [], []
                for c0 in range(b, e, 2048):
                    c1 = min(c0 + 2048, e)
                    z = O.z_vals_uniform(torch.full((c1 - c0, 1), fr['near']), torch.full((c1 - c0, 1), fr['far']), N_SAMPLES)
                    h, p = O.render_head_torso_chunk(sd, ro[c0:c1], rd[c0:c1], rot[c0:c1], rdt[c0:c1], z, fr['bc_rgb'][c0:c1],
                                                     zs, za, sig, sig_t)
                    hs.append(h)
                    ps.append(p)
                return {'rgb_head': torch.cat(hs), 'rgb_map': torch.cat(ps)}
    else:
        n_imp = 0 if workload == 'coarse64' else N_IMPORTANCE
        sd_c, sd_f = synth.facenerf_state_dict(0), synth.facenerf_state_dict(1)
        if use_ref:
            model = ref_arm.FaceNeRFFrame(sd_c, sd_f if n_imp else None, N_SAMPLES, n_imp)

            def run():
                return model.render(hw, hw, fr['focal'], fr['cx'], fr['cy'], fr['c2w'], fr['bc_rgb'], fr['aud'], fr['near'],
                                    fr['far'], (b, e))
        else:
            def run():
                return O.render(hw, hw, fr['focal'], fr['cx'], fr['cy'], fr['c2w'], fr['bc_rgb'], fr['aud'], sd_c, sd_f,
                                fr['near'], fr['far'], N_SAMPLES, n_imp, chunk=2048, ray_slice=(b, e),
                                keys=('rgb_map', 'rgb0', 'z_samples') if n_imp else ('rgb_map',))
    times, out = [], None
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            out = run()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return dict(value=rays_per_step / med, sec=med, cores=torch.get_num_threads(), kind='reference' if use_ref else 'port',
                slice=(b, e), out=out, rays=rays_per_step)


WORKLOADS = {
    'facenerf': 'FaceNeRF 450x450 x (64 coarse + 128 fine samples), synthetic seeded weights/pose/latent '
                '(BASELINE.json configs[1]); one step = one frame',
    'head_torso': 'Decoder head + torso (DeformationField_ori) two-field frame, 450x450 x 64 samples, synthetic seeded '
                  'weights/poses/latents (BASELINE.json configs[2], the path scripts/test_obama.sh runs); one step = one frame',
    'coarse64': 'FaceNeRF single-frame render, 64x64 rays x 64 coarse samples, no fine pass (BASELINE.json configs[0], the '
                'reference\'s own CPU-sized case); one step = one frame',
    'mlp_1m': 'Synthetic random-weight 8x256 NeRF (no latent), 2^20 rays x 192 samples, flat network query + compositing '
              '(BASELINE.json configs[3]); one step = all rays',
    'sequence': 'Audio-driven FaceNeRF sequence, 450x450 x (64+128) per frame, per-frame pose + latent tables, uint8 frames '
                'copied out double-buffered; several GPUs: every frame sharded by rays, uint8 tiles gathered on a side stream '
                '(BASELINE.json configs[4]); one step = the sequence',
}
WORKLOADS['train_step'] = 'Training step of the live model (MAIN:764-931): 2048 random rays x 64 samples x (head + torso) Decoder fields, ' \
                          'two-field compositing, two MSE losses, backward, Adam on decoder + AudNet + ExpNet; synthetic 450x450 targets; ' \
                          'one step = one optimiser step (single GPU: the reference trains on one)'
EVALS = {'facenerf': '(64+192) FaceNeRF', 'head_torso': '(64 head + 64 torso) Decoder', 'coarse64': '64 FaceNeRF'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    wl = args.workload if args.workload in EVALS else 'facenerf'
    r = cpu_arm(wl, max(1, args.steps), max(1, min(args.warmup, 1)), CPU_RAYS)
    sample = '%d rays (chunks of 2048, image centre) x %s evaluations per ray per step, median of %d' % (
        r['rays'], EVALS[wl], max(1, args.steps))
    hw = 64 if wl == 'coarse64' else H
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': r['value'], 'unit': 'rays/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': r['sec'] * 1e3, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOADS[wl] + '; CPU arm renders a bounded ray sample of the same frame',
                   'rays_per_frame': hw * hw, 'rays_per_step': r['rays']},
        'cpu_baseline': {'value': r['value'], 'unit': 'rays/s', 'cores': r['cores'], 'kind': r['kind'], 'sample': sample},
        'e2e': {'value': r['value'], 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


# --------------------------------------------------------------------------------------------- GPU arm
class Job:
    """One workload at one precision on this rank: the device-resident step, the end-to-end step, and what to report."""

    def __init__(self, workload, precision, args, ctx):
        import torch
        import dfa_nerf_b200 as dfn
        from dfa_nerf_b200.distributed import shard_range
        import synth                       # seeded synthetic data (repo root; nothing under oracle/ is touched on this path)
        self.workload, self.precision, self.ctx, self.dfn, self.torch = workload, precision, ctx, dfn, torch
        dev, rank, world = ctx['dev'], ctx['rank'], ctx['world']
        self.prec = {'bf16': dfn.PREC_BF16, 'fp16': dfn.PREC_FP16, 'bf16x3': dfn.PREC_BF16X3, 'fp16x3m': dfn.PREC_FP16X3M, 'fp32': dfn.PREC_FP32}[precision]
        prec = self.prec
        hw = 64 if workload == 'coarse64' else H
        self.hw = hw
        fr = self.fr = synth.frame_inputs(H=hw, W=hw, seed=0)
        self.n_rays = hw * hw
        self.b, self.e, _ = shard_range(self.n_rays, rank, world)
        self.launches = 0
        self.sinks = {}
        self.frames = args.frames
        tc_name = ('mlp_pp_kernel<%s>' % precision) if prec in (dfn.PREC_BF16X3, dfn.PREC_FP16X3M) else 'mlp_pair_kernel<%s>' % precision
        self.flops_note = 'algorithmic, latent/viewdir columns folded: 2*557,184 per MLP evaluation (BASELINE.md section 2)'

        def mk(seed):
            m = dfn.FaceNeRF(D=8, W=256, input_ch=63, input_ch_views=27, dim_aud=64, output_ch=4, skips=[4], use_viewdirs=True)
            m.load_state_dict(synth.facenerf_state_dict(seed))
            return m.to(dev)

        if workload in ('facenerf', 'coarse64'):
            n_imp = N_IMPORTANCE if workload == 'facenerf' else 0
            self.eng = dfn.RenderEngine(mk(0), mk(1) if n_imp else None, N_SAMPLES, n_imp, precision=prec)
            self.bc_dev, self.bc_host = fr['bc_rgb'].to(dev), fr['bc_rgb'].pin_memory()
            self.aud_dev, self.lat_host = fr['aud'].to(dev), fr['aud'].pin_memory()
            self.evals_per_ray = N_SAMPLES + (N_SAMPLES + n_imp if n_imp else 0)
            self.kernel_name = tc_name
        elif workload == 'mlp_1m':
            self.n_rays = 1 << 20
            self.b, self.e, _ = shard_range(self.n_rays, rank, world)
            S = N_SAMPLES + N_IMPORTANCE
            self.net = dfn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=4, skips=[4], use_viewdirs=True)
            self.net.load_state_dict(synth.nerf_state_dict(0))
            self.net = self.net.to(dev)
            self.eng = dfn.RenderEngine(self.net, None, S, 0, precision=prec)
            g = torch.Generator().manual_seed(0)
            n_loc = self.e - self.b
            self.ro = fr['c2w'][:3, -1].expand(n_loc, 3).contiguous().to(dev)
            self.rd = (torch.randn(n_loc, 3, generator=g) * 0.2 + torch.tensor([0., 0., -1.])).to(dev)
            self.vd = self.rd / torch.norm(self.rd, dim=-1, keepdim=True)
            z, _ = torch.sort(torch.rand(n_loc, S, generator=g) * 0.6 + 0.4, -1)
            self.z = z.to(dev)
            self.bc_dev = torch.rand(self.n_rays, 3, generator=g).to(dev)
            self.bc_host = self.bc_dev.cpu().pin_memory()
            self.aud_dev, self.lat_host = torch.zeros(1, device=dev), torch.zeros(1).pin_memory()
            self.evals_per_ray = S
            self.kernel_name = tc_name
            self.flops_note = 'algorithmic, viewdir columns folded: 2*557,184 per MLP evaluation (NeRF and FaceNeRF coincide, SURVEY 8d)'
        elif workload == 'train_step':
            from dfa_nerf_b200.train import Trainer, select_coords
            import numpy as np
            if world > 1:
                raise SystemExit('--workload train_step is single-GPU (the reference trains on one GPU; no gradient exchange is built)')
            dec = dfn.Decoder(z_dim=256, hidden_size=256, dim_signal=96, use_deformation_field=True)
            dec.load_state_dict(synth.decoder_state_dict(0))
            aud, exp = dfn.AudioNet_W2L(), dfn.ExpressionEnc()
            aud.load_state_dict(synth.mlp_encoder_state_dict(7))
            exp.load_state_dict(synth.mlp_encoder_state_dict(8, (64, 32, 32)))
            tprec = dfn.PREC_BF16X3 if precision in ('bf16x3', 'fp32') else dfn.PREC_BF16
            self.trainer = Trainer(dec, aud, exp, lrate=5e-4, N_samples=N_SAMPLES, precision=tprec, device=dev)
            g = torch.Generator().manual_seed(3)
            poses = torch.cat([synth.pose_sequence(8, 0), torch.tensor([0., 0., 0., 1.]).expand(8, 1, 4)], 1)
            self.host_imgs = [torch.rand(H, W, 3, generator=g).pin_memory() for _ in range(3)]     # target_com, target_head_neck, bc
            np.random.seed(0)
            self.batch = dict(H=H, W=W, focal=fr['focal'], cx=fr['cx'], cy=fr['cy'], near=fr['near'], far=fr['far'], poses=poses, img_i=3,
                              pose=poses[3, :3, :4], pose_torso=poses[0, :3, :4], auds=torch.randn(8, 512, generator=g).to(dev),
                              exps=torch.randn(8, 64, generator=g).to(dev), coords=select_coords(H, W, [100, 125, 200, 200], 2048, 0.95).to(dev),
                              target_com=self.host_imgs[0].to(dev), target_head_neck=self.host_imgs[1].to(dev), bc_img=self.host_imgs[2].to(dev),
                              z_shape=torch.randn(1, 2, 256, generator=g).to(dev), z_app=torch.randn(1, 2, 256, generator=g).to(dev))
            self.n_rays, self.b, self.e = 2048, 0, 2048
            self.evals_per_ray = 2 * N_SAMPLES
            self.kernel_name = 'gemm_tc_kernel<%s>' % ('bf16x3' if tprec == dfn.PREC_BF16X3 else 'bf16')
            self.flops_note = 'algorithmic 2*M*N*K of every GEMM launch of the step (forward, data and weight gradients, per-frame ' \
                              'vector products); bf16x3 issues three MMAs per product and counts once'
            self.gstep = 0
            self.lat_host = torch.zeros(1).pin_memory()
        elif workload == 'sequence':
            self.eng = dfn.RenderEngine(mk(0), mk(1), N_SAMPLES, N_IMPORTANCE, precision=prec)
            self.seq = synth.frame_inputs(H=H, W=W, seed=0, n_frames=args.frames)
            self.evals_per_ray = N_SAMPLES + N_SAMPLES + N_IMPORTANCE
            self.kernel_name = tc_name
            self.n_rays = H * W * args.frames            # rays per step: the whole sequence
            self.b, self.e = 0, self.n_rays
            self.lat_host = self.seq['aud'].pin_memory()
            self.bc_dev = fr['bc_rgb'].to(dev)
            self.aud_dev = None
            # several GPUs: every frame split by rays (1/G of the frame latency, no tail imbalance; DFN_BENCH_SEQ_SHARD=frames: whole
            # frames per rank, the round-1 form)
            self.seq_shard = os.environ.get('DFN_BENCH_SEQ_SHARD', 'rays') if world > 1 else 'frames'
            # the sequence's pinned output buffer belongs to the caller (pinning 182 MB is ~0.1 s: not a per-sequence cost of a job
            # that renders sequence after sequence)
            self.seq_out = torch.empty((args.frames, H, W, 3), dtype=torch.uint8).pin_memory() \
                if (rank == 0 and (world == 1 or self.seq_shard == 'rays')) else None
        else:
            if prec == dfn.PREC_FP32:
                raise SystemExit('--workload head_torso runs on the tensor-core path: --precision bf16 | fp16 | bf16x3')
            self.dec = dfn.Decoder(z_dim=256, hidden_size=256, dim_signal=96, use_deformation_field=True)
            self.dec.load_state_dict(synth.decoder_state_dict(0))
            self.dec = self.dec.to(dev)
            self.c2w_torso = synth.frame_inputs(H=H, W=W, seed=7)['c2w']        # fixed body pose (MAIN:644)
            zs, za, sig, sig_t = _head_torso_latents()
            self.zs, self.za = zs.to(dev), za.to(dev)
            lat = torch.cat([sig.reshape(-1), sig_t.reshape(-1)])              # head signal | torso signal
            self.aud_dev, self.lat_host = lat.to(dev), lat.pin_memory()
            self.bc_dev, self.bc_host = fr['bc_rgb'].to(dev), fr['bc_rgb'].pin_memory()
            self.evals_per_ray = 2 * N_SAMPLES
            self.kernel_name = ('mlp_pp_kernel<%s, Decoder>' if precision == 'bf16x3' else 'mlp_pair_kernel<%s, Decoder>') % precision
            self.flops_note = 'algorithmic, per-frame latents and per-ray view term folded: 2*556,032 (head) + 2*628,352 (torso ' \
                              'incl. deformation field) per sample (SURVEY.md section 8d, appendix A)'

    # -- one frame on this rank's ray range
    def render(self, bc_full, lat):
        dfn, fr, b, e = self.dfn, self.fr, self.b, self.e
        if self.workload in ('facenerf', 'coarse64'):
            out = self.eng.render_frame(self.hw, self.hw, fr['focal'], fr['c2w'], bc_full, lat, fr['near'], fr['far'], fr['cx'], fr['cy'],
                                        ray_range=(b, e), want=('rgb_map',))
            self.launches += self.eng.last_launches + 1            # + get_rays
            return out['rgb_map']
        if self.workload == 'mlp_1m':
            raw = self.eng.query_points(self.net, self.ro, self.rd, self.vd, self.z, None)
            self.launches += self.eng.last_launches + 1
            return dfn.raw2outputs(raw, self.z, self.rd, bc_full[b:e])[0]
        _, rgb = dfn.render_head_torso(self.dec, H, W, fr['focal'], fr['c2w'], self.c2w_torso, bc_full, self.zs, self.za, lat[:96],
                                       lat[96:], fr['near'], fr['far'], fr['cx'], fr['cy'], N_samples=N_SAMPLES, ray_range=(b, e),
                                       precision=self.prec)
        self.launches += dfn.render_head_torso.last_launches
        return rgb

    def step_sequence(self, n=None):
        # pose / latent tables from (pinned) host memory, uint8 frames back to pinned host memory: this IS the end-to-end path
        seq, world = self.seq, self.ctx['world']
        n = self.frames if n is None else n
        frames = self.dfn.render_sequence(self.eng, H, W, seq['focal'], seq['c2w_seq'][:n], self.lat_host[:n], self.bc_dev, seq['near'],
                                          seq['far'], seq['cx'], seq['cy'], shard=self.seq_shard, out=self.seq_out)
        self.launches += (self.eng.last_launches + 2) * (n if self.seq_shard == 'rays' else (n + world - 1) // world)
        return frames

    def step_train(self, e2e=False):
        b = self.batch
        if e2e:     # the step's images come from the host (the reference reads three JPEGs per step, MAIN:771-774); the loss goes back
            b = dict(b, target_com=self.host_imgs[0].to(self.ctx['dev'], non_blocking=True),
                     target_head_neck=self.host_imgs[1].to(self.ctx['dev'], non_blocking=True),
                     bc_img=self.host_imgs[2].to(self.ctx['dev'], non_blocking=True))
        loss = self.trainer.step(b, global_step=self.gstep)
        self.gstep += 1
        self.launches += self.trainer.last_launches
        return float(loss) if e2e else loss

    def sink(self, to_host):
        if to_host not in self.sinks:
            self.sinks[to_host] = self.dfn.RayShardSink(self.n_rays, self.ctx['dev'], to_host=to_host, depth=3)
        return self.sinks[to_host]

    def step_resident(self):
        if self.workload == 'train_step':
            return self.step_train()
        if self.workload == 'sequence':
            return self.step_sequence()
        rgb = self.render(self.bc_dev, self.aud_dev)
        if self.ctx['world'] == 1:
            return rgb
        return self.sink(False).push(rgb)       # the all-gather of frame i runs on the sink's stream under frame i+1's kernels

    def step_e2e(self):
        torch, dev = self.torch, self.ctx['dev']
        if self.workload == 'train_step':
            return self.step_train(e2e=True)
        if self.workload == 'sequence':
            return self.step_sequence()
        bc = self.bc_host[self.b:self.e].to(dev, non_blocking=True)
        lat = self.lat_host.to(dev, non_blocking=True)
        bc_full = torch.empty((self.n_rays, 3), dtype=torch.float32, device=dev)
        bc_full[self.b:self.e] = bc
        rgb = self.render(bc_full, lat)
        # frame loop as a user of the package writes it (dfa_nerf_b200.RayShardSink, three buffers): frame i's all-gather and rank 0's
        # copy into pinned host memory run on a side stream while the next frames render; every step the host takes frame i-2 (blocking
        # on ITS copy), i.e. it runs at most two frames ahead of the device -- with eight ranks coupled through the gather, one frame of
        # slack left the GPUs waiting for the slowest host of the previous frame
        sink = self.sink(True)
        i = sink.push(rgb)
        if i > 1:
            sink.wait(i - 2)

    def points_per_step(self):
        """Network evaluations this rank performs per step."""
        if self.workload == 'sequence':
            world, rank = self.ctx['world'], self.ctx['rank']
            if self.seq_shard == 'rays':
                from dfa_nerf_b200.distributed import shard_range
                b, e, _ = shard_range(H * W, rank, world)
                return (e - b) * self.evals_per_ray * self.frames
            f0, f1 = self.dfn.shard_frames(self.frames, rank, world)
            return H * W * self.evals_per_ray * (f1 - f0)
        return (self.e - self.b) * self.evals_per_ray

    def bytes_per_step(self):
        if self.workload == 'train_step':
            return int(3 * H * W * 3 * 4), 4
        if self.workload == 'sequence':
            return int(self.lat_host.numel() * 4 + self.frames * 48), int(self.n_rays * 3)
        return int((self.e - self.b) * 12 + self.lat_host.numel() * 4 + 48), (int(self.n_rays * 12) if self.ctx['rank'] == 0 else 0)


def timed(ctx, fn, steps, warmup):
    import torch
    import torch.distributed as dist
    world, dev = ctx['world'], ctx['dev']
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


def measure(ctx, job, steps, warmup, e2e_warmup=2):
    """Device-resident throughput with per-launch CUDA events around the tcgen05 kernel (recorded by the library on the stream
    it launches on: dfn_profile_enable / collect), then the end-to-end step.  Returns a dict of the reported numbers."""
    import torch
    dfn = job.dfn
    for _ in range(warmup):
        if job.workload == 'sequence':
            job.step_sequence(min(job.frames, 2 * ctx['world']))     # warm-up on a short prefix of the sequence
        else:
            job.step_resident()
    torch.cuda.synchronize()
    dfn.lib.dfn_profile_enable(1 if job.prec != dfn.PREC_FP32 else 0)
    job.launches = 0
    ms_step = timed(ctx, job.step_resident, steps, 0)
    k_ms, k_n, k_macs = C.c_double(), C.c_int64(), C.c_double()
    dfn.lib.dfn_profile_collect(C.byref(k_ms), C.byref(k_n), C.byref(k_macs))
    dfn.lib.dfn_profile_enable(0)
    n_launch = job.launches
    ms_e2e = ms_step if job.workload == 'sequence' else timed(ctx, job.step_e2e, steps, e2e_warmup)   # the sequence step IS end to end
    h2d, d2h = job.bytes_per_step()
    res = {'value': job.n_rays / (ms_step * 1e-3), 'ms_per_step': ms_step, 'gpu_launches': n_launch,
           'e2e': {'value': job.n_rays / (ms_e2e * 1e-3), 'unit': 'rays/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
           'roofline': None}
    pk = peaks()
    if k_n.value > 0:
        achieved = 2.0 * k_macs.value / (k_ms.value * 1e-3) / 1e12
        # points per launch, averaged over the launches of the timed region: ncu's per-point DRAM / L2->SM bytes scale with it
        pts_per_launch = job.points_per_step() * steps / k_n.value
        tr = ncu_traffic().get(job.kernel_name)
        res['roofline'] = {'bound': 'tensor', 'achieved': achieved, 'peak': pk['bf16_sustained'], 'unit': 'TFLOP/s',
                           'frac': achieved / pk['bf16_sustained'],
                           'traffic': tr['dram_bytes_per_point'] * pts_per_launch if tr else None,
                           'l2_to_sm_bytes': tr['l2_to_sm_bytes_per_point'] * pts_per_launch if tr else None,
                           'traffic_source': tr['source'] if tr else None,
                           'algorithmic_bytes': 16.0 * pts_per_launch,
                           'kernel': job.kernel_name, 'launches': int(k_n.value), 'avg_launch_ms': k_ms.value / k_n.value,
                           'kernel_share_of_step': k_ms.value / (ms_step * steps), 'flops': job.flops_note,
                           'peak_source': 'bf16 dense sustained, ' + pk['source']}
    return res


def parity_leg(ctx, cpu, workload, precisions):
    """max-abs RGB error of the GPU path, per precision, against the CPU arm's render of ITS ray range (the checker role of
    oracle/: nothing here is timed).  FaceNeRF: teacher-forced (the CPU arm's z_samples injected into the fine pass, SURVEY
    section 7) and free-running; head_torso has no resampling, so the one number is both."""
    import torch
    import dfa_nerf_b200 as dfn
    import synth
    dev = ctx['dev']
    b, e = cpu['slice']
    ref = cpu['out']
    fr = synth.frame_inputs(H=H, W=W, seed=0)
    out = {'rays': e - b, 'against': 'cpu_baseline (kind %s), rays [%d, %d) of the frame' % (cpu['kind'], b, e)}
    err = lambda a, r: (a.detach().cpu().double() - r.double()).abs()                                  # noqa: E731
    if workload == 'head_torso':
        dec = dfn.Decoder(z_dim=256, hidden_size=256, dim_signal=96, use_deformation_field=True)
        dec.load_state_dict(synth.decoder_state_dict(0))
        dec = dec.to(dev)
        zs, za, sig, sig_t = [t.to(dev) for t in _head_torso_latents()]
        c2w_t = synth.frame_inputs(H=H, W=W, seed=7)['c2w']
        for name in precisions:
            prec = {'bf16': dfn.PREC_BF16, 'fp16': dfn.PREC_FP16, 'bf16x3': dfn.PREC_BF16X3}[name]
            _, p = dfn.render_head_torso(dec, H, W, fr['focal'], fr['c2w'], c2w_t, fr['bc_rgb'].to(dev), zs, za, sig, sig_t, fr['near'],
                                         fr['far'], fr['cx'], fr['cy'], N_samples=N_SAMPLES, ray_range=(b, e), precision=prec)
            d = err(p, ref['rgb_map'])
            out[name] = {'max_abs_rgb': d.max().item(), 'p99': d.flatten().quantile(0.99).item(), 'median': d.median().item()}
        return out

    def mk(seed):
        m = dfn.FaceNeRF(D=8, W=256, input_ch=63, input_ch_views=27, dim_aud=64, output_ch=4, skips=[4], use_viewdirs=True)
        m.load_state_dict(synth.facenerf_state_dict(seed))
        return m.to(dev)
    nc, nf = mk(0), mk(1)
    ro, rd, vd = [t.reshape(-1, 3)[b:e].contiguous() for t in dfn.get_rays(H, W, fr['focal'], fr['c2w'], fr['cx'], fr['cy'], device=dev,
                                                                            return_viewdirs=True)]
    near, far = torch.full((e - b,), fr['near'], device=dev), torch.full((e - b,), fr['far'], device=dev)
    bc, aud = fr['bc_rgb'][b:e].to(dev), fr['aud'].to(dev)
    for name in precisions:
        prec = {'bf16': dfn.PREC_BF16, 'fp16': dfn.PREC_FP16, 'bf16x3': dfn.PREC_BF16X3, 'fp16x3m': dfn.PREC_FP16X3M, 'fp32': dfn.PREC_FP32}[name]
        eng = dfn.RenderEngine(nc, nf, N_SAMPLES, N_IMPORTANCE, precision=prec)
        tf = eng.render_rays(ro, rd, vd, near, far, bc, aud, z_samples=ref['z_samples'].to(dev), want=('rgb_map', 'rgb0'))
        free = eng.render_rays(ro, rd, vd, near, far, bc, aud, want=('rgb_map',))
        d, dfree = err(tf['rgb_map'], ref['rgb_map']), err(free['rgb_map'], ref['rgb_map'])
        out[name] = {'teacher_forced_max_abs_rgb': d.max().item(), 'teacher_forced_p99': d.flatten().quantile(0.99).item(),
                     'coarse_max_abs_rgb': err(tf['rgb0'], ref['rgb0']).max().item(),
                     'free_running_max_abs_rgb': dfree.max().item(), 'free_running_p99': dfree.flatten().quantile(0.99).item(),
                     'free_running_median': dfree.median().item(), 'meets_1e-4': bool(d.max().item() <= 1e-4)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default='bf16', choices=list(PRECISIONS))
    ap.add_argument('--workload', default='facenerf', choices=sorted(WORKLOADS))
    ap.add_argument('--frames', type=int, default=300, help='sequence workload: frames per step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip modes / parity / extra (the other configs)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import dfa_nerf_b200 as dfn      # noqa: F401  (fails loudly without libdfn.so)

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a GPU (there is no CPU path)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    ctx = {'world': world, 'rank': rank, 'dev': dev}

    # ---- the headline workload: clocks are sampled from its warm-up through both of its timed regions
    job = Job(args.workload, args.precision, args, ctx)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    sampler.mark()
    main_res = measure(ctx, job, args.steps, args.warmup)
    t_main_end = time.time()

    extras = not args.no_extras
    modes, extra = {}, {}
    if extras and args.workload in ('facenerf', 'head_torso'):
        for name in ('bf16', 'fp16', 'bf16x3') + (('fp16x3m',) if args.workload == 'facenerf' else ()):
            if name == args.precision:
                continue
            r = measure(ctx, Job(args.workload, name, args, ctx), max(3, min(args.steps, 5)), 3)
            modes[name] = {'value': r['value'], 'ms_per_step': r['ms_per_step'], 'e2e': r['e2e']['value'],
                           'frac': r['roofline']['frac'] if r['roofline'] else None,
                           'kernel': r['roofline']['kernel'] if r['roofline'] else None}
    if extras and args.workload == 'facenerf':
        plan = [('head_torso', 'head_torso', max(3, min(args.steps, 5))), ('coarse64', 'coarse64', 10), ('mlp_1m', 'mlp_1m', 3),
                ('sequence', 'sequence_%d' % args.frames, 1)]
        if world == 1:
            plan.append(('train_step', 'train_step', 10))
        for wl, key, steps in plan:
            j = Job(wl, 'bf16x3' if wl == 'train_step' else args.precision, args, ctx)
            r = measure(ctx, j, steps, 3)
            extra[key] = {'workload': WORKLOADS[wl], 'value': r['value'], 'unit': 'rays/s', 'ms_per_step': r['ms_per_step'],
                          'rays_per_step': j.n_rays, 'e2e': r['e2e'], 'gpu_launches': r['gpu_launches'], 'steps': steps,
                          'frac': r['roofline']['frac'] if r['roofline'] else None,
                          'kernel_share_of_step': r['roofline']['kernel_share_of_step'] if r['roofline'] else None}
            if wl == 'train_step':
                extra[key]['precision'] = 'bf16x3'
                extra[key]['points_per_step'] = j.points_per_step()
            if wl == 'sequence':
                extra[key]['frames'] = args.frames
                extra[key]['ms_per_frame'] = r['ms_per_step'] / args.frames
            del j
    clocks = sampler.stop(t_main_end)

    cpu, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload in EVALS:
        c = cpu_arm(args.workload, 3, 1, CPU_RAYS)
        cpu = {'value': c['value'], 'unit': 'rays/s', 'cores': c['cores'], 'kind': c['kind'],
               'sample': '%d rays (chunks of 2048, image centre) x %s network evaluations per ray, median of 3 (%.1f s each)'
                         % (c['rays'], EVALS[args.workload], c['sec'])}
        if extras and args.workload in ('facenerf', 'head_torso'):
            parity = parity_leg(ctx, c, args.workload, [p for p in ('bf16', 'fp16', 'bf16x3') + (('fp16x3m',) if args.workload == 'facenerf' else ())])
            parity['precision'] = args.precision
            if args.workload == 'facenerf':
                parity['teacher_forced_max_abs_rgb'] = parity[args.precision]['teacher_forced_max_abs_rgb'] \
                    if args.precision in parity else None
        if extras and args.workload == 'facenerf' and 'coarse64' in extra:
            c0 = cpu_arm('coarse64', 1, 1, 64 * 64)
            extra['coarse64']['cpu_baseline'] = {'value': c0['value'], 'unit': 'rays/s', 'cores': c0['cores'], 'kind': c0['kind'],
                                                 'sample': 'the whole 64x64 frame, one run (%.1f s)' % c0['sec']}

    if rank == 0:
        print(json.dumps({
            'metric': METRIC, 'value': main_res['value'], 'unit': 'rays/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': main_res['ms_per_step'], 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': {'bf16': 'bf16', 'fp16': 'fp16', 'bf16x3': 'bf16x3 (split bf16, fp32-parity)',
                      'fp16x3m': 'fp16x3m (fp16, split on the density layers: fp32-parity)', 'fp32': 'f32'}[args.precision],
            'data': 'synthetic',
            'config': {'workload': WORKLOADS[args.workload],
                       'rays_per_step': job.n_rays, 'mlp_evals_per_ray': job.evals_per_ray, 'precision': args.precision,
                       'parallelism': ('every frame ray-sharded over %d GPU(s), uint8 tiles all-gathered per frame on a side stream' % world)
                       if args.workload == 'sequence' else 'single GPU' if args.workload == 'train_step' else 'rays sharded over %d GPU(s), one all-gather of the RGB tile per frame on a side stream (frame i gathers / copies out while frame i+1 renders)' % world,
                       'l2': 'per-step intermediates (~1.4 GB of raw/z buffers) exceed the 126 MB L2; no explicit flush'},
            'e2e': main_res['e2e'], 'gpu_launches': main_res['gpu_launches'], 'clocks': clocks, 'roofline': main_res['roofline'],
            'cpu_baseline': cpu, 'modes': modes or None, 'parity': parity, 'extra': extra or None,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

"""Turns gpurun_out/ ncu artefacts into the committed text summaries under profiles/.
    python profiles/summarize.py r01
Reads gpurun_out/launches_<round>.csv (ncu --metrics gpu__time_duration.sum of `bench.py --steps 2`) and
gpurun_out/prof_<round>_*.ncu-rep (ncu --set full) with `ncu -i ... --page raw --csv`."""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else 'r01'
out = []


def launches(suffix='', cmd='bench.py --steps 2 --warmup 3'):
    p = os.path.join(ROOT, 'gpurun_out', 'launches_%s%s.csv' % (rnd, suffix))
    if not os.path.exists(p):
        return
    rows = list(csv.DictReader(l for l in open(p) if l.startswith('"')))
    agg = collections.OrderedDict()
    for r in rows:
        k = r['Kernel Name'].split('(')[0].replace('void ', '')
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r['Metric Value'])
    tot = sum(v[1] for v in agg.values())
    out.append('## Launch list (%s): ncu --metrics gpu__time_duration.sum --clock-control none, `%s`' % (rnd, cmd))
    out.append('(cold-cache, serialised launches: compare SHARES, not absolutes; %d launches, %.1f ms total)\n' % (len(rows), tot / 1e6))
    out.append('%-52s %6s %12s %12s %8s' % ('kernel', 'n', 'total ms', 'avg us', 'share'))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append('%-52s %6d %12.3f %12.2f %8.4f' % (k[:52], n, t / 1e6, t / n / 1e3, t / tot))
    out.append('')


WANT = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__cluster_size', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__sass_inst_executed_op_tmem_ldt.sum', 'sm__inst_executed_pipe_tensor_subpipe_hmma.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']


def report(tag, title):
    p = os.path.join(ROOT, 'gpurun_out', 'prof_%s_%s.ncu-rep' % (rnd, tag))
    if not os.path.exists(p):
        return
    txt = subprocess.run(['ncu', '-i', p, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out.append('## %s  (ncu --set full --clock-control none; %s)' % (title, os.path.basename(p)))
    for vals in rows[2:]:
        d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        out.append('kernel: %s   grid %s block %s' % (d.get('Kernel Name', ('', '?'))[1], d.get('Grid Size', ('', '?'))[1], d.get('Block Size', ('', '?'))[1]))
        for k in WANT:
            if k in d and d[k][1] != '':
                out.append('  %-82s %12s %s' % (k, d[k][1], d[k][0]))
        try:
            rd = float(d['dram__bytes_read.sum'][1].replace(',', ''))
            wr = float(d['dram__bytes_write.sum'][1].replace(',', ''))
            mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
            out.append('  DRAM traffic per launch: %.3f MB' % ((rd * mult[d['dram__bytes_read.sum'][0]] + wr * mult[d['dram__bytes_write.sum'][0]]) / 1e6))
        except Exception:
            pass
    out.append('')


TRAFFIC = {}


def traffic(tag, kernel_name, points_per_launch, also=()):
    """Per-point DRAM and L2->SM (TMA) bytes of a tcgen05 kernel from its ncu --set full capture of the bench's own launches
    -> profiles/traffic.json (read by bench.py for roofline.traffic / l2_to_sm_bytes)."""
    p = os.path.join(ROOT, 'gpurun_out', 'prof_%s_%s.ncu-rep' % (rnd, tag))
    if not os.path.exists(p):
        return
    txt = subprocess.run(['ncu', '-i', p, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
    dram = l2 = 0.0
    for vals in rows[2:]:
        d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        g = lambda k: float(d[k][1].replace(',', '')) * mult[d[k][0]]          # noqa: E731
        dram += g('dram__bytes_read.sum') + g('dram__bytes_write.sum')
        l2 += g('l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum')
    pts = float(sum(points_per_launch))
    rec = {'dram_bytes_per_point': dram / pts, 'l2_to_sm_bytes_per_point': l2 / pts,
           'source': 'ncu --set full --clock-control none of the bench\'s own launches (profiles/ncu_summary_%s.txt, %s): %d launches, '
                     '%.0f points' % (rnd, os.path.basename(p), len(rows) - 2, pts)}
    TRAFFIC[kernel_name] = rec
    for other in also:
        TRAFFIC[other] = dict(rec, source=rec['source'] + '; same kernel schedule and data movement as ' + kernel_name)


if rnd == 'r01':
    launches()
    launches('_ht', 'bench.py --workload head_torso --steps 2 --warmup 3')
    report('bf16', 'mlp_tc_kernel<bf16>: 20,000 rays x 192 samples (fine-pass sized network query)')
    report('x3', 'mlp_pp_kernel<bf16x3>: 20,000 rays x 192 samples')
    report('dec_all', 'mlp_pp_kernel<bf16, Decoder>: 60,000 rays x 64 samples, head field then torso field (live model)')
    report('vw', 'volume_weights_kernel<1> (raw2outputs): 202,500 rays x 64 samples (coarse), then x 192 (fine)')
    report('headtorso', 'head_torso_kernel (live two-field compositing): 202,500 rays x 64 samples')
    report('sortmerge', 'sort_merge_kernel: 202,500 rays x (64 + 128)')
    report('samplepdf', 'sample_pdf_kernel: 202,500 rays, 63 bins -> 128 samples')
elif rnd == 'r02b':
    # late round 2: register-resident render_prep_kernel, early staging / folded density heads in the split kernels, RayShardSink
    frame = 202500
    launches('', 'bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline')
    launches('_ht', 'bench.py --workload head_torso --steps 2 --warmup 3 --no-extras --no-cpu-baseline')
    launches('_train', 'bench.py --workload train_step --steps 2 --warmup 3 --no-extras --no-cpu-baseline (twelve-warp GEMM readout)')
    report('bench_pair', 'mlp_pair_kernel<bf16> (per-ray bias rows of the view layer staged in shared memory): the bench frame\'s coarse (202,500 x 64) and fine (202,500 x 192) launches')
    report('bench_x3m', 'mlp_pp_kernel<fp16x3m> (early staging, alpha_linear in the last trunk layer\'s epilogue): the bench frame\'s coarse and fine launches')
    report('bench_x3', 'mlp_pp_kernel<bf16x3> (early staging): the bench frame\'s coarse and fine launches')
    report('bench_dec_x3', 'mlp_pp_kernel<bf16x3, Decoder> (folded-head programs, sigma_out in fp32 in the epilogue): the head_torso frame\'s head and torso launches')
    report('stages', 'HBM-bound stage kernels of the bench frame: render_prep_kernel<27>, coarse_to_fine_kernel<2>, volume_weights_kernel<1,6>')
    traffic('bench_pair', 'mlp_pair_kernel<bf16>', [frame * 64, frame * 192], also=('mlp_pair_kernel<fp16>',))
    traffic('bench_x3m', 'mlp_pp_kernel<fp16x3m>', [frame * 64, frame * 192])
    traffic('bench_x3', 'mlp_pp_kernel<bf16x3>', [frame * 64, frame * 192])
    traffic('bench_dec_x3', 'mlp_pp_kernel<bf16x3, Decoder>', [frame * 64, frame * 64])
    if TRAFFIC:
        import json
        tp = os.path.join(ROOT, 'profiles', 'traffic.json')
        old = json.load(open(tp)) if os.path.exists(tp) else {}
        old.update(TRAFFIC)
        json.dump(old, open(tp, 'w'), indent=1)
else:
    launches('', 'bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline')
    launches('_ht', 'bench.py --workload head_torso --steps 2 --warmup 3 --no-extras --no-cpu-baseline')
    launches('_fused', 'bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline (fused stage kernels)')
    launches('_train', 'bench.py --workload train_step --steps 2 --warmup 3 --no-extras --no-cpu-baseline')
    frame = 202500
    report('bench_pair', 'mlp_pair_kernel<bf16> (CTA pairs, cta_group::2; the default since this round): the bench frame\'s coarse (202,500 x 64) and fine (202,500 x 192) launches')
    report('bench_bf16', 'mlp_tc_kernel<bf16> (1-CTA generation, the default until this round): the bench frame\'s coarse (202,500 x 64) and fine (202,500 x 192) launches')
    report('bench_x3', 'mlp_pp_kernel<bf16x3>: the bench frame\'s coarse and fine launches')
    report('bench_dec', 'mlp_pair_kernel<bf16> running the Decoder programs: the head_torso frame\'s head and torso launches (202,500 x 64 each)')
    report('bench_dec_x3', 'mlp_pp_kernel<bf16x3, Decoder>: the head_torso frame\'s head and torso launches in the split precision')
    report('embed', 'embed_kernel (dfn_embed, HELP:21-52): 12.96 M points -> [P,63] (12 B in + 252 B out per point)')
    report('stages', 'HBM-bound stage kernels of the bench frame')
    report('gemm', 'gemm_tc_kernel (training step GEMMs)')
    traffic('bench_pair', 'mlp_pair_kernel<bf16>', [frame * 64, frame * 192], also=('mlp_pair_kernel<fp16>',))
    traffic('bench_bf16', 'mlp_tc_kernel<bf16>', [frame * 64, frame * 192], also=('mlp_tc_kernel<fp16>',))
    traffic('bench_x3', 'mlp_pp_kernel<bf16x3>', [frame * 64, frame * 192])
    traffic('bench_dec', 'mlp_pair_kernel<bf16, Decoder>', [frame * 64, frame * 64], also=('mlp_pair_kernel<fp16, Decoder>',))
    traffic('bench_dec_x3', 'mlp_pp_kernel<bf16x3, Decoder>', [frame * 64, frame * 64])
    if TRAFFIC:
        import json
        old = {}
        tp = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(tp):
            old = json.load(open(tp))
        old.update(TRAFFIC)
        json.dump(old, open(tp, 'w'), indent=1)
path = os.path.join(ROOT, 'profiles', 'ncu_summary_%s.txt' % rnd)
open(path, 'w').write('\n'.join(out) + '\n')
print('\n'.join(out))

"""CPU: the reference arm of bench.py (`--impl reference`: the reference's own code from oracle/_ref, or the oracle port, on the host
cores) prints one JSON line with
the keys the driver's contract names; the GPU arm refuses to run without a device."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('workload,kind', [('facenerf', 'reference'), ('head_torso', 'reference'), ('coarse64', 'port')])
def test_reference_arm_json_line(workload, kind):
    """kind 'reference' = the reference's own modules vendored into oracle/_ref by oracle/build_ref.py (present wherever
    build() ran with /root/reference mounted -- it travels to the GPU box); 'port' = the restatement in oracle/nerf_oracle.py."""
    sys.path.insert(0, ROOT)
    from oracle import ref_arm
    if kind == 'reference' and not ref_arm.available():
        kind = 'port'
    env = dict(os.environ, DFN_BENCH_CPU_RAYS='192', OMP_NUM_THREADS='4', DFN_BENCH_CPU_KIND=kind)
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', workload,
                          '--steps', '1', '--warmup', '1'], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d['impl'] == 'reference' and d['unit'] == 'rays/s' and d['higher_is_better'] is True
    assert d['metric'].startswith('rendered rays/sec at 450x450x(64+128)')
    assert d['value'] > 0 and d['ms_per_step'] > 0 and d['steps'] == 1
    assert d['cpu_baseline']['kind'] == kind and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['config']['rays_per_step'] == 192 and 'workload' in d['config'] and d['vs_baseline'] is None


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU behaviour')
def test_gpu_arm_fails_loudly_without_a_device():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1'], capture_output=True, text=True, timeout=600)
    assert out.returncode != 0 and 'needs a GPU' in out.stderr

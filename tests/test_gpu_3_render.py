"""GPU parity of the hierarchical render_rays pipeline through dfn_render_rays.
Gates: stage-wise / teacher-forced <= 1e-4; free-running end to end is reported
(max / p99 / median) because coarse->fine resampling is ill-conditioned."""
import pytest
import torch

from oracle import nerf_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def maxerr(a, b):
    return (a.detach().cpu().double() - b.detach().cpu().double()).abs().max().item()


@pytest.fixture(scope='module')
def dfn():
    import dfa_nerf_b200
    return dfa_nerf_b200


def nets(dfn, cs, fs):
    out = []
    for s in (cs, fs):
        m = dfn.FaceNeRF(D=8, W=256, input_ch=63, input_ch_views=27, dim_aud=64, output_ch=4, skips=[4], use_viewdirs=True)
        m.load_state_dict(synth.facenerf_state_dict(s))
        out.append(m.to(DEV))
    return out


def golden_rays(g):
    ro, rd = O.get_rays(g['H'], g['W'], g['focal'], g['c2w'], g['cx'], g['cy'])
    ro, rd = ro.reshape(-1, 3).contiguous(), rd.reshape(-1, 3)
    vd = rd / torch.norm(rd, dim=-1, keepdim=True)
    n = ro.shape[0]
    return ro, rd, vd, torch.full((n,), g['near']), torch.full((n,), g['far'])


@pytest.mark.parametrize('prec_name,tol', [('PREC_FP32', 2e-5), ('PREC_BF16X3', 1e-4), ('PREC_FP16X3M', 1e-4)])
def test_render_rays_golden_teacher_forced(dfn, golden, prec_name, tol):
    g = golden('render_rays')
    nc, nf = nets(dfn, g['coarse_seed'], g['fine_seed'])
    ro, rd, vd, near, far = golden_rays(g)
    eng = dfn.RenderEngine(nc, nf, 64, 128, precision=getattr(dfn, prec_name))
    out = eng.render_rays(ro.to(DEV), rd.to(DEV), vd.to(DEV), near.to(DEV), far.to(DEV), g['bc_rgb'].to(DEV),
                          g['aud'].to(DEV), z_samples=g['z_samples'].to(DEV),
                          want=('rgb_map', 'disp_map', 'acc_map', 'last_weight', 'rgb0', 'z_vals'))
    assert torch.equal(out['z_vals'].cpu(), g['z_vals'])          # merged sample depths: bit-exact
    assert maxerr(out['rgb0'], g['rgb0']) < tol
    assert maxerr(out['rgb_map'], g['rgb_map']) < tol
    assert maxerr(out['acc_map'], g['acc_map']) < tol
    assert maxerr(out['last_weight'], g['weights'][:, -1]) < tol
    assert torch.allclose(out['disp_map'].cpu(), g['disp_map'], rtol=1e-3)


def test_render_rays_free_running_report(dfn, golden):
    g = golden('render_rays')
    nc, nf = nets(dfn, g['coarse_seed'], g['fine_seed'])
    ro, rd, vd, near, far = golden_rays(g)
    for prec_name in ('PREC_FP32', 'PREC_BF16X3', 'PREC_FP16X3M', 'PREC_FP16', 'PREC_BF16'):
        eng = dfn.RenderEngine(nc, nf, 64, 128, precision=getattr(dfn, prec_name))
        out = eng.render_rays(ro.to(DEV), rd.to(DEV), vd.to(DEV), near.to(DEV), far.to(DEV), g['bc_rgb'].to(DEV),
                              g['aud'].to(DEV), want=('rgb_map', 'z_samples'))
        e = (out['rgb_map'].cpu() - g['rgb_map']).abs().reshape(-1)
        zs = (out['z_samples'].cpu() - g['z_samples']).abs().max().item()
        print('%s free-running: rgb max %.2e p99 %.2e median %.2e | z_samples max %.2e'
              % (prec_name, e.max(), e.kthvalue(int(0.99 * e.numel())).values, e.median(), zs))
        assert torch.isfinite(out['rgb_map']).all()
        if prec_name in ('PREC_FP32', 'PREC_BF16X3', 'PREC_FP16X3M'):
            assert e.median() < 1e-5 and e.max() < 5e-2
        elif prec_name == 'PREC_FP16':
            assert e.median() < 2e-4
        else:
            assert e.median() < 5e-3


def test_upstream_render_signature(dfn, golden):
    """render(H, W, focal, cx, cy, chunk, c2w=..., **render_kwargs) as run_nerf.py calls it; ragged chunks."""
    g = golden('render_rays')
    nc, nf = nets(dfn, g['coarse_seed'], g['fine_seed'])
    kw = dict(network_fn=nc, network_fine=nf, N_samples=64, N_importance=128, perturb=0., white_bkgd=False,
              raw_noise_std=0., network_query_fn=None, precision=dfn.PREC_FP32)
    rgb, disp, acc, last_w, extras = dfn.render(g['H'], g['W'], g['focal'], g['cx'], g['cy'], chunk=100,
                                                c2w=g['c2w'].to(DEV), bc_rgb=g['bc_rgb'].to(DEV), aud_para=g['aud'].to(DEV),
                                                near=g['near'], far=g['far'], use_viewdirs=True, **kw)
    assert rgb.shape == (g['H'], g['W'], 3) and disp.shape == (g['H'], g['W']) and 'rgb0' in extras
    assert maxerr(extras['rgb0'].reshape(-1, 3), g['rgb0']) < 2e-5
    e = (rgb.reshape(-1, 3).cpu() - g['rgb_map']).abs()
    assert e.median() < 1e-5


def test_coarse_only_matches_live_config(dfn, golden):
    """N_importance=0 with 64 uniform samples: the sampling the reference's live path uses (MAIN:617-619)."""
    g = golden('render_rays')
    nc, _ = nets(dfn, g['coarse_seed'], g['fine_seed'])
    ro, rd, vd, near, far = golden_rays(g)
    eng = dfn.RenderEngine(nc, None, 64, 0, precision=dfn.PREC_BF16X3)
    out = eng.render_rays(ro.to(DEV), rd.to(DEV), vd.to(DEV), near.to(DEV), far.to(DEV), g['bc_rgb'].to(DEV),
                          g['aud'].to(DEV), want=('rgb_map', 'last_weight'))
    assert maxerr(out['rgb_map'], g['rgb0']) < 1e-4
    assert maxerr(out['last_weight'], g['weights0'][:, -1]) < 1e-4


def test_full_frame_properties(dfn):
    """450x450x(64+128), the BASELINE.json size: size-independent properties instead of a CPU oracle
    (a full frame is minutes on the CPU): determinism, chunk invariance, bounded colours,
    acc + last_weight consistency, and bf16x3 == fp32 kernels on a random ray subset."""
    fr = synth.frame_inputs(H=450, W=450, seed=0)
    nc, nf = nets(dfn, 0, 1)
    bc, aud = fr['bc_rgb'].to(DEV), fr['aud'].to(DEV)
    eng = dfn.RenderEngine(nc, nf, 64, 128, precision=dfn.PREC_BF16X3)
    full = eng.render_frame(450, 450, fr['focal'], fr['c2w'], bc, aud, fr['near'], fr['far'], fr['cx'], fr['cy'],
                            want=('rgb_map', 'acc_map', 'last_weight'))
    rgb = full['rgb_map']
    assert rgb.shape == (202500, 3) and torch.isfinite(rgb).all()
    assert rgb.min() >= -1e-5 and rgb.max() <= 1 + 1e-5
    again = eng.render_frame(450, 450, fr['focal'], fr['c2w'], bc, aud, fr['near'], fr['far'], fr['cx'], fr['cy'])
    assert torch.equal(again['rgb_map'], rgb)                               # deterministic
    part = eng.render_frame(450, 450, fr['focal'], fr['c2w'], bc, aud, fr['near'], fr['far'], fr['cx'], fr['cy'],
                            ray_range=(100000, 103333))
    assert torch.equal(part['rgb_map'], rgb[100000:103333])                 # ray independence / tail tiles
    assert (full['acc_map'] - 1).abs().max() < 1e-3                          # last sample absorbs the rest
    e32 = dfn.RenderEngine(nc, nf, 64, 128, precision=dfn.PREC_FP32)
    sub = e32.render_frame(450, 450, fr['focal'], fr['c2w'], bc, aud, fr['near'], fr['far'], fr['cx'], fr['cy'],
                           ray_range=(50000, 54096), want=('rgb_map', 'rgb0'))
    sub3 = eng.render_frame(450, 450, fr['focal'], fr['c2w'], bc, aud, fr['near'], fr['far'], fr['cx'], fr['cy'],
                            ray_range=(50000, 54096), want=('rgb_map', 'rgb0'))
    assert maxerr(sub['rgb0'], sub3['rgb0']) < 1e-4
    e = (sub['rgb_map'] - sub3['rgb_map']).abs()
    assert e.median() < 1e-5


def test_to8b_bit_exact(dfn):
    """HELP:17 on the device against numpy's (255*clip(x,0,1)).astype(uint8), ragged length, edge values."""
    g = torch.Generator().manual_seed(0)
    x = torch.cat([torch.rand(4099, generator=g) * 1.4 - 0.2, torch.tensor([0., 1., -0., 0.999999, 1 / 255., 254.5 / 255., 2., -3.])])
    ref = torch.from_numpy(O.to8b(x.numpy()))
    assert torch.equal(dfn.to8b(x.to(DEV)).cpu(), ref)
    y = torch.rand(5, 7, 3, generator=g)
    assert torch.equal(dfn.to8b(y.to(DEV)).cpu(), torch.from_numpy(O.to8b(y.numpy())))


def test_render_sequence_matches_per_frame(dfn):
    """Sequence loop (device-side to8b + double-buffered D2H) == frame-by-frame render + to8b, FaceNeRF and live model."""
    H = W = 12
    n = 5
    fr = synth.frame_inputs(H=H, W=W, seed=2, n_frames=n)
    net_c, net_f = nets(dfn, 0, 1)
    eng = dfn.RenderEngine(net_c, net_f, 64, 128, precision=dfn.PREC_BF16X3)
    bc = fr['bc_rgb'].to(DEV)
    frames = dfn.render_sequence(eng, H, W, fr['focal'], fr['c2w_seq'], fr['aud'], bc, fr['near'], fr['far'], fr['cx'], fr['cy'])
    assert frames.shape == (n, H, W, 3) and frames.dtype == torch.uint8 and not frames.is_cuda
    for i in range(n):
        one = eng.render_frame(H, W, fr['focal'], fr['c2w_seq'][i], bc, fr['aud'][i].to(DEV), fr['near'], fr['far'], fr['cx'], fr['cy'])
        assert torch.equal(frames[i], dfn.to8b(one['rgb_map']).reshape(H, W, 3).cpu())
    dec = dfn.Decoder(z_dim=256, hidden_size=256, dim_signal=96, use_deformation_field=True)
    dec.load_state_dict(synth.decoder_state_dict(1))
    dec = dec.to(DEV)
    g = torch.Generator().manual_seed(5)
    zs, za = torch.randn(1, 2, 256, generator=g).to(DEV), torch.randn(1, 2, 256, generator=g).to(DEV)
    sig, sig_t = torch.randn(n, 96, generator=g), torch.randn(n, 42, generator=g)
    body = synth.camera_pose(9)
    fr2 = dfn.render_sequence_head_torso(dec, H, W, fr['focal'], fr['c2w_seq'], body, bc, zs, za, sig, sig_t, fr['near'], fr['far'],
                                         fr['cx'], fr['cy'])
    for i in (0, n - 1):
        _, person = dfn.render_head_torso(dec, H, W, fr['focal'], fr['c2w_seq'][i], body, bc, zs, za, sig[i].to(DEV), sig_t[i].to(DEV),
                                          fr['near'], fr['far'], fr['cx'], fr['cy'])
        assert torch.equal(fr2[i], dfn.to8b(person).reshape(H, W, 3).cpu())


def test_render_rays_maximum_samples_and_limits(dfn):
    """The largest supported ray (128 coarse + 128 fine = 256 samples, eight per lane in the compositing kernels, the
    512-wide padded merge) teacher-forced against the oracle; one sample more is refused with an error, not truncated."""
    R, Nc, Nf = 24, 128, 128
    fr = synth.frame_inputs(H=R, W=1, seed=6)
    sd_c, sd_f = synth.facenerf_state_dict(0), synth.facenerf_state_dict(1)
    nc, nf = nets(dfn, 0, 1)
    ro, rd = O.get_rays(R, 1, fr['focal'], fr['c2w'], fr['cx'], fr['cy'])
    ro, rd = ro.reshape(-1, 3).contiguous(), rd.reshape(-1, 3)
    vd = rd / torch.norm(rd, dim=-1, keepdim=True)
    near, far = torch.full((R,), fr['near']), torch.full((R,), fr['far'])
    rays = torch.cat([ro, rd, near[:, None], far[:, None], vd], -1)
    with torch.no_grad():
        ref = O.render_rays(rays, fr['bc_rgb'], fr['aud'], sd_c, sd_f, Nc, Nf)
    eng = dfn.RenderEngine(nc, nf, Nc, Nf, precision=dfn.PREC_BF16X3)
    out = eng.render_rays(ro.to(DEV), rd.to(DEV), vd.to(DEV), near.to(DEV), far.to(DEV), fr['bc_rgb'].to(DEV), fr['aud'].to(DEV),
                          z_samples=ref['z_samples'].to(DEV), want=('rgb_map', 'rgb0', 'acc_map', 'z_vals'))
    assert torch.equal(out['z_vals'].cpu(), ref['z_vals'])
    assert maxerr(out['rgb0'], ref['rgb0']) < 1e-4 and maxerr(out['rgb_map'], ref['rgb_map']) < 1e-4
    assert maxerr(out['acc_map'], ref['acc_map']) < 1e-4
    with pytest.raises(dfn.DfnError):
        dfn.RenderEngine(nc, nf, 129, 128, precision=dfn.PREC_BF16X3).render_rays(
            ro.to(DEV), rd.to(DEV), vd.to(DEV), near.to(DEV), far.to(DEV), fr['bc_rgb'].to(DEV), fr['aud'].to(DEV))


def test_render_rays_ragged_ray_count(dfn):
    """A ray count that is no multiple of the preparation kernel's 8-ray blocks (render_prep_kernel: per-ray view-bias rows and
    coarse depths of both networks), teacher-forced against the oracle; the last rays are the ones a block-tail bug would miss."""
    R = 37
    fr = synth.frame_inputs(H=R, W=1, seed=11)
    sd_c, sd_f = synth.facenerf_state_dict(0), synth.facenerf_state_dict(1)
    nc, nf = nets(dfn, 0, 1)
    ro, rd = O.get_rays(R, 1, fr['focal'], fr['c2w'], fr['cx'], fr['cy'])
    ro, rd = ro.reshape(-1, 3).contiguous(), rd.reshape(-1, 3)
    vd = rd / torch.norm(rd, dim=-1, keepdim=True)
    near, far = torch.full((R,), fr['near']), torch.full((R,), fr['far'])
    rays = torch.cat([ro, rd, near[:, None], far[:, None], vd], -1)
    with torch.no_grad():
        ref = O.render_rays(rays, fr['bc_rgb'], fr['aud'], sd_c, sd_f, 64, 128)
    eng = dfn.RenderEngine(nc, nf, 64, 128, precision=dfn.PREC_BF16X3)
    out = eng.render_rays(ro.to(DEV), rd.to(DEV), vd.to(DEV), near.to(DEV), far.to(DEV), fr['bc_rgb'].to(DEV), fr['aud'].to(DEV),
                          z_samples=ref['z_samples'].to(DEV), want=('rgb_map', 'rgb0', 'z_vals'))
    assert torch.equal(out['z_vals'].cpu(), ref['z_vals'])
    assert maxerr(out['rgb0'], ref['rgb0']) < 1e-4 and maxerr(out['rgb_map'], ref['rgb_map']) < 1e-4
    assert maxerr(out['rgb_map'][-5:], ref['rgb_map'][-5:]) < 1e-4


def test_ray_shard_sink_and_sequence_out_buffer(dfn):
    """The pipelined output side on the device (one process: no gather): RayShardSink's side-stream copies with three buffers and two
    frames of host slack deliver every frame intact and in order, also into a caller-owned pinned buffer; render_sequence(out=...) fills
    that buffer and returns views of it."""
    H = W = 12
    n = 7
    fr = synth.frame_inputs(H=H, W=W, seed=3, n_frames=n)
    net_c, net_f = nets(dfn, 0, 1)
    eng = dfn.RenderEngine(net_c, net_f, 64, 128, precision=dfn.PREC_BF16X3)
    bc = fr['bc_rgb'].to(DEV)
    ref = [eng.render_frame(H, W, fr['focal'], fr['c2w_seq'][i], bc, fr['aud'][i].to(DEV), fr['near'], fr['far'], fr['cx'], fr['cy'])['rgb_map'].clone()
           for i in range(n)]
    sink = dfn.RayShardSink(H * W, torch.device(DEV), depth=3)
    seq = torch.empty((n, H * W, 3), dtype=torch.float32).pin_memory()
    got = {}
    for i in range(n):
        tile = eng.render_frame(H, W, fr['focal'], fr['c2w_seq'][i], bc, fr['aud'][i].to(DEV), fr['near'], fr['far'], fr['cx'], fr['cy'])['rgb_map']
        k = sink.push(tile, host_out=seq[i] if i % 2 else None)       # ring slots and caller-owned rows alternate
        assert k == i
        if i > 1:
            got[i - 2] = sink.wait(i - 2).clone()
    for i in (n - 2, n - 1):
        got[i] = sink.wait(i).clone()
    sink.finish()
    for i in range(n):
        assert torch.equal(got[i], ref[i].cpu()), i
        if i % 2:
            assert torch.equal(seq[i], ref[i].cpu())
    with pytest.raises(IndexError):
        sink.wait(0)
    out = torch.empty((n, H, W, 3), dtype=torch.uint8).pin_memory()
    frames = dfn.render_sequence(eng, H, W, fr['focal'], fr['c2w_seq'], fr['aud'], bc, fr['near'], fr['far'], fr['cx'], fr['cy'], out=out)
    assert frames.data_ptr() == out.data_ptr()
    for i in range(n):
        assert torch.equal(frames[i], dfn.to8b(ref[i]).reshape(H, W, 3).cpu())
    with pytest.raises(dfn.DfnError):
        dfn.render_sequence(eng, H, W, fr['focal'], fr['c2w_seq'], fr['aud'], bc, fr['near'], fr['far'], fr['cx'], fr['cy'],
                            out=torch.empty((n, H, W, 3), dtype=torch.uint8))          # not pinned

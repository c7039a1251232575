"""GPU parity of the stage kernels (through the C ABI) against the oracle and the golden vectors.
Bars: rays / depths / merged depths / searchsorted indices bit-exact; floats to the stated tolerance."""
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


def test_get_rays_bit_exact(dfn, golden):
    g = golden('get_rays')
    for tag in 'abc':
        H, W, f, cx, cy, stride = [float(v) for v in g[tag + '_args']]
        o, d = dfn.get_rays(int(H), int(W), f, g[tag + '_c2w'], None if cx < 0 else cx, None if cy < 0 else cy, int(stride))
        assert torch.equal(o.cpu(), g[tag + '_o']), tag
        assert torch.equal(d.cpu(), g[tag + '_d']), tag
    # full 450x450 frame against the oracle
    c2w = synth.camera_pose(3)
    o, d, v = dfn.get_rays(450, 450, 1200., c2w.to(DEV), 225., 225., return_viewdirs=True)
    ro, rd = O.get_rays(450, 450, 1200., c2w, 225., 225.)
    assert torch.equal(d.cpu(), rd) and torch.equal(o.cpu(), ro.contiguous())
    assert maxerr(v, rd / torch.norm(rd, dim=-1, keepdim=True)) < 2e-7
    assert torch.equal(d.cpu().reshape(-1, 3)[g['d_idx']], g['d_d'])


def test_z_vals_bit_exact(dfn, golden):
    g = golden('z_vals')
    near = torch.tensor([0.4, 0.3, 0.55], device=DEV)
    far = torch.tensor([1.0, 0.9, 1.7], device=DEV)
    z = dfn.z_vals_uniform(near, far, 64)
    ref = O.z_vals_uniform(near.cpu()[:, None], far.cpu()[:, None], 64)
    assert torch.equal(z.cpu(), ref)
    assert torch.equal(z.cpu()[0], g['z64'])
    rnd = torch.rand(3, 64)
    zs = dfn.z_vals_uniform(near, far, 64, perturb_rand=rnd.to(DEV))
    assert torch.equal(zs.cpu(), O.z_vals_stratified(ref, rnd))


def test_embed(dfn, golden):
    g = golden('embed')
    x = g['x'].to(DEV)
    for L, key in ((10, 'pe10'), (4, 'pe4'), (3, 'pe3')):
        fn, dim = dfn.get_embedder(L, 0)
        y = fn(x)
        assert y.shape == (96, dim) and dim == 3 + 6 * L
        assert maxerr(y, g[key]) < 1e-6          # sin/cos: CUDA libm vs CPU libm, <= 2 ulp on [-1,1]
        assert torch.equal(y[:, :3].cpu(), g['x'])
    assert maxerr(dfn.decoder_transform_points(x[None], 10), g['tp10'][None]) < 1e-6
    assert maxerr(dfn.decoder_transform_points(x[None], 4, views=True), g['tp4'][None]) < 1e-6
    # seeded, and checked against the float64 restatement so a miss can only be the kernel's
    big = (torch.rand(200000, 3, generator=torch.Generator().manual_seed(11)) * 2 - 1) * 1.2
    assert maxerr(dfn.get_embedder(10)[0](big.to(DEV)), O.embed(big.double(), 10).float()) < 1e-6


def test_calc_volume_weights(dfn, golden):
    g = golden('composite')
    w = dfn.calc_volume_weights(g['z'][None].to(DEV), g['rays_d'][None].to(DEV), g['sigma'][None].to(DEV))
    assert w.shape == (1, 24, 64)
    assert maxerr(w[0], g['weights']) < 1e-6
    # 192 samples, ragged ray count
    R, S = 1000, 192
    z, _ = torch.sort(torch.rand(R, S) * 0.6 + 0.4, -1)
    rd = torch.randn(R, 3)
    sg = torch.randn(R, S) * 12 + 2
    w = dfn.calc_volume_weights(z.to(DEV), rd.to(DEV), sg.to(DEV))
    assert maxerr(w, O.calc_volume_weights(z, rd, sg)) < 2e-6    # expf: 2 ulp (CUDA) vs 1 ulp (host libm) on values <= 1


def test_composite_function(dfn, golden):
    g = golden('composite')
    ss, fw = dfn.composite_function(g['sigma2'][:, None].to(DEV), g['feat2'][:, None].to(DEV))
    assert ss.shape == (1, 24, 64) and fw.shape == (1, 24, 64, 3)
    assert torch.equal(ss[0].cpu(), g['sigma_sum'])
    assert maxerr(fw[0], g['feat_w']) == 0.0
    ss1, fw1 = dfn.composite_function(g['sigma2'][:1, None].to(DEV), g['feat2'][:1, None].to(DEV))
    assert torch.equal(ss1[0].cpu(), g['sigma2'][0]) and torch.equal(fw1[0].cpu(), g['feat2'][0])


def test_raw2outputs(dfn, golden):
    g = golden('raw2outputs')
    rgb, disp, acc, w, depth = dfn.raw2outputs(g['raw'].to(DEV), g['z'].to(DEV), g['rays_d'].to(DEV), g['bc_rgb'].to(DEV))
    assert maxerr(rgb, g['rgb_map']) < 1e-6
    assert maxerr(acc, g['acc_map']) < 1e-6
    assert maxerr(w, g['weights']) < 1e-6
    assert maxerr(depth, g['depth_map']) < 1e-6
    assert torch.allclose(disp.cpu(), g['disp_map'], rtol=1e-5)
    # white background + 192 samples
    R, S = 777, 192
    raw = torch.randn(R, S, 4)
    raw[..., 3] = raw[..., 3] * 10 + 1
    z, _ = torch.sort(torch.rand(R, S) * 0.6 + 0.4, -1)
    rd = torch.randn(R, 3)
    bc = torch.rand(R, 3)
    out = dfn.raw2outputs(raw.to(DEV), z.to(DEV), rd.to(DEV), bc.to(DEV), white_bkgd=True)
    ref = O.raw2outputs(raw, z, rd, bc, white_bkgd=True)
    for k in (0, 2, 3, 4):                       # rgb, acc, weights, depth
        assert maxerr(out[k], ref[k]) < 2e-6
    assert torch.allclose(out[1].cpu(), ref[1], rtol=1e-4)


def test_sample_pdf_indices_and_samples(dfn, golden):
    g = golden('sample_pdf')
    bins, wts = g['bins'].to(DEV), g['weights'].to(DEV)
    # (1) inversion on the oracle's own cdf: indices and samples bit-exact
    w = g['weights'] + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cat([torch.zeros(24, 1), torch.cumsum(pdf, -1)], -1)
    u = O.linspace_table(128)
    s, i = dfn.invert_cdf(bins, cdf.to(DEV), u.to(DEV))
    assert torch.equal(i.cpu(), g['det_inds'])
    assert torch.equal(s.cpu(), g['det'])
    s, i = dfn.invert_cdf(bins, cdf.to(DEV), g['u_py'].to(DEV))
    assert torch.equal(i.cpu(), g['py_inds'])
    assert torch.equal(s.cpu(), g['py'])
    # (2) full sample_pdf.  The kernel builds the cdf with an exact (fp64) row sum and torch's fp64 running
    # sum; torch's own fp32 row sum is vectorised differently per CPU ISA, so against the plain oracle a
    # sample may differ only where u sits within rounding of a cdf knot (always possible at u == 1.0, where
    # the denom<1e-5 switch (HELP:577) can move the sample by one bin).  Against the oracle evaluated with
    # the same exact row sum, indices and samples must be identical.
    s, i = dfn.sample_pdf(bins, wts, 128, det=True, return_inds=True)
    es, ei = exact_sum_oracle(g['bins'], g['weights'], O.linspace_table(128))
    assert torch.equal(i.cpu(), ei) and torch.equal(s.cpu(), es)
    check_against_plain_oracle(s, i, g['det'], g['det_inds'])
    s2, i2 = dfn.sample_pdf(bins, wts, 128, u=g['u_py'].to(DEV), return_inds=True)
    es, ei = exact_sum_oracle(g['bins'], g['weights'], g['u_py'])
    assert torch.equal(i2.cpu(), ei) and torch.equal(s2.cpu(), es)
    check_against_plain_oracle(s2, i2, g['py'], g['py_inds'])
    # reference's pytest switch (HELP:552-561)
    s3 = dfn.sample_pdf(bins, wts, 128, det=False, pytest=True)
    assert torch.equal(s3, s2)
    # edge cases: all-zero weights, one-hot weights, tiny weights
    s4, i4 = dfn.sample_pdf(bins[:3], g['edge_weights'].to(DEV), 128, det=True, return_inds=True)
    es, ei = exact_sum_oracle(g['bins'][:3], g['edge_weights'], O.linspace_table(128))
    assert torch.equal(i4.cpu(), ei) and torch.equal(s4.cpu(), es)
    check_against_plain_oracle(s4, i4, g['edge'], g['edge_inds'])
    assert torch.equal(s4[:, 0].cpu(), g['bins'][:3, 0])


def exact_sum_oracle(bins, weights, u):
    """oracle sample_pdf with the row sum taken in fp64 (the only step whose fp32 order torch leaves open)."""
    w = weights + 1e-5
    pdf = w / w.double().sum(-1, keepdim=True).float()
    cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
    u = u.expand(cdf.shape[0], u.shape[-1]).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below, above = torch.clamp(inds - 1, min=0), torch.clamp(inds, max=cdf.shape[-1] - 1)
    cb, ca = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bb, ba = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    den = ca - cb
    den = torch.where(den < 1e-5, torch.ones_like(den), den)
    return bb + (u - cb) / den * (ba - bb), inds


def check_against_plain_oracle(s, i, ref_s, ref_i):
    s, i = s.cpu(), i.cpu()
    bad_i = i != ref_i
    bad_s = (s - ref_s).abs() > 1e-6
    assert bad_i.float().mean().item() < 0.01 and bad_s.float().mean().item() < 0.01
    # samples that keep their index agree to the conditioning of the lerp: the two cdf knots of a bin each differ by up
    # to one ulp between the two row-sum orders, which moves the sample by at most 2 ulp/denom * bin width with
    # denom >= 1e-5 (HELP:577) -> 1.2e-7/1e-5 * 0.0095 = 1.1e-4
    assert ((s - ref_s).abs()[~bad_i]).max().item() < 2e-4
    assert ((s - ref_s).abs()[~bad_i]).median().item() < 1e-7


def test_sample_pdf_large(dfn):
    R = 5000
    z = O.z_vals_uniform(torch.full((R, 1), 0.4), torch.ones(R, 1), 64).expand(R, 64).contiguous()
    zm = .5 * (z[..., 1:] + z[..., :-1])
    w = torch.rand(R, 62) ** 8
    s, i = dfn.sample_pdf(zm.to(DEV), w.to(DEV), 128, det=True, return_inds=True)
    es, ei = exact_sum_oracle(zm, w, O.linspace_table(128))
    assert torch.equal(i.cpu(), ei) and torch.equal(s.cpu(), es)
    rs, ri = O.sample_pdf(zm, w, 128, det=True, return_inds=True)
    check_against_plain_oracle(s, i, rs, ri)
    assert torch.all(s[:, 1:-1] >= s[:, :-2])          # det samples are sorted


def test_sort_merge_bit_exact(dfn):
    R = 3001
    a, _ = torch.sort(torch.rand(R, 64), -1)
    b = torch.rand(R, 128)
    out = dfn.sort_merge(a.to(DEV), b.to(DEV))
    ref, _ = torch.sort(torch.cat([a, b], -1), -1)
    assert torch.equal(out.cpu(), ref)
    out = dfn.sort_merge(a[:5, :7].contiguous().to(DEV), b[:5, :3].contiguous().to(DEV))
    assert torch.equal(out.cpu(), torch.sort(torch.cat([a[:5, :7], b[:5, :3]], -1), -1)[0])
    # render-time case: both runs ascending (merge-by-rank fast path), with ties inside and across the runs, mixed per
    # ray with unsorted rows (bitonic fallback), and the largest supported width
    bs, _ = torch.sort(b, -1)
    bs[::3, 5] = bs[::3, 4]
    bs[::5, :64:7] = a[::5, :64:7]
    bs, _ = torch.sort(bs, -1)
    mixed = bs.clone()
    mixed[1::2] = b[1::2]
    for bb in (bs, mixed):
        assert torch.equal(dfn.sort_merge(a.to(DEV), bb.to(DEV)).cpu(), torch.sort(torch.cat([a, bb], -1), -1)[0])
    big_a, big_b = torch.sort(torch.rand(33, 500), -1)[0], torch.sort(torch.rand(33, 524), -1)[0]
    assert torch.equal(dfn.sort_merge(big_a.to(DEV), big_b.to(DEV)).cpu(), torch.sort(torch.cat([big_a, big_b], -1), -1)[0])


@pytest.mark.parametrize('R,Nc,Nf', [(1, 64, 128), (777, 64, 128), (50, 16, 16), (33, 100, 60), (20000, 64, 128)])
def test_coarse_to_fine_bit_identical_to_the_stage_chain(dfn, R, Nc, Nf):
    """The fused coarse -> fine kernel (dfn_coarse_to_fine: raw2outputs -> z_mid -> sample_pdf -> sort-merge in one launch,
    intermediates in shared memory) against the chain of the separately tested stage kernels: same operations in the same
    order, so the coarse image, the new samples and the merged depths must be bit-identical -- for the deterministic u table
    (merge-by-rank path), for random per-ray u (unsorted samples: bitonic path) and with injected samples (teacher forcing)."""
    g = torch.Generator().manual_seed(R + Nc)
    raw = torch.randn(R, Nc, 4, generator=g)
    raw[..., 3] = raw[..., 3] * 12 + 2
    raw[: max(1, R // 10), :, 3] = -1.                         # empty rays: uniform pdf
    z0, _ = torch.sort(torch.rand(R, Nc, generator=g) * 0.6 + 0.4, -1)
    rd = torch.randn(R, 3, generator=g) * 0.2 + torch.tensor([0., 0., -1.])
    bc = torch.rand(R, 3, generator=g)
    raw, z0, rd, bc = [t.to(DEV) for t in (raw, z0, rd, bc)]

    def chain(u, zs_in=None):
        rgb0, _, _, w, _ = dfn.raw2outputs(raw, z0, rd, bc)
        if zs_in is None:
            zmid = (0.5 * (z0[:, 1:] + z0[:, :-1])).contiguous()
            zs = dfn.sample_pdf(zmid, w[:, 1:-1].contiguous(), Nf, u=u)
        else:
            zs = zs_in
        return dfn.sort_merge(z0, zs), zs, rgb0

    u_det = torch.linspace(0., 1., Nf).to(DEV)
    u_rand = torch.rand(R, Nf, generator=g).to(DEV)
    for u in (u_det, u_rand):
        za, zs, rgb0 = dfn.coarse_to_fine(raw, z0, rd, Nf, bc, u=u)
        ra, rs, r0 = chain(u)
        assert torch.equal(zs, rs) and torch.equal(za, ra) and torch.equal(rgb0, r0)
        assert (za[:, 1:] >= za[:, :-1]).all()
    inj = (torch.rand(R, Nf, generator=g) * 0.6 + 0.4).to(DEV)
    za, zs, rgb0 = dfn.coarse_to_fine(raw, z0, rd, Nf, bc, z_samples=inj)
    ra, rs, r0 = chain(None, inj)
    assert torch.equal(za, ra) and torch.equal(zs, inj) and torch.equal(rgb0, r0)
    za2, _, none = dfn.coarse_to_fine(raw, z0, rd, Nf, bc, u=u_det, want_rgb0=False)
    assert none is None and torch.equal(za2, chain(u_det)[0])

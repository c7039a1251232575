"""GPU: the `--render_person` job end to end (MAIN:590-733) from a preprocessed-identity directory to JPEG files:
loader (LOAD:14-47) -> latent pre-pass -> head+torso frame loop -> to8b -> FrameWriter, against the oracle chain."""
import os

import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = 'cuda'
FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'audface_tiny')


def test_render_person_directory_to_jpegs(tmp_path):
    import dfa_nerf_b200 as dfn
    from PIL import Image
    data = dfn.load_audface_data_split(FIX, testskip=1, test_file='transforms_val_ba.json', aud_file='aud.pt')
    body = dfn.pose_body(FIX, use_ba=True)
    sd_a, sd_e, sd_d = synth.audionet_state_dict(4, dim_aud=64), synth.mlp_encoder_state_dict(2, (64, 32, 32)), synth.decoder_state_dict(2)
    a, e = dfn.AudioNet(dim_aud=64, win_size=16), dfn.ExpressionEnc()
    dec = dfn.Decoder(z_dim=256, hidden_size=256, dim_signal=96, use_deformation_field=True)
    for m, sd in ((a, sd_a), (e, sd_e), (dec, sd_d)):
        m.load_state_dict(sd)
        m.to(DEV)
    g = torch.Generator().manual_seed(5)
    zs, za = torch.randn(1, 2, 256, generator=g), torch.randn(1, 2, 256, generator=g)
    near, far = 0.3, 0.9
    seen = []
    paths = dfn.render_person(dec, data, body, zs.to(DEV), za.to(DEV), a, e, str(tmp_path), near, far, N_samples=64,
                              precision=dfn.PREC_BF16X3, video=str(tmp_path / 'person.mp4'))
    n = data['poses'].shape[0]
    assert [os.path.basename(p) for p in paths] == ['test_%06d.jpg' % i for i in range(n)]
    assert sorted(os.listdir(tmp_path / 'render_head')) == [os.path.basename(p) for p in paths]
    assert os.path.getsize(tmp_path / 'person.mp4') > 0

    H, W, focal, cx, cy = data['hwfcxy']
    bc = torch.from_numpy(np.array(data['bc_img'])).float().reshape(-1, 3) / 255.0
    poses, auds, exps = [torch.from_numpy(data[k]) for k in ('poses', 'auds', 'exp')]
    rot, rdt = [t.reshape(-1, 3) for t in O.get_rays(H, W, focal, body[:3, :4], cx, cy)]
    z = O.z_vals_uniform(torch.full((H * W, 1), near), torch.full((H * W, 1), far), 64)
    with torch.no_grad():
        for i in range(n):
            sig = torch.cat([O.audionet_forward(sd_a, auds[i:i + 1]), O.mlp_encoder_forward(sd_e, exps[i:i + 1])], 1)
            sig_t = O.encode_signal_torso(poses, i)
            ro, rd = [t.reshape(-1, 3) for t in O.get_rays(H, W, focal, poses[i, :3, :4], cx, cy)]
            head, person = O.render_head_torso_chunk(sd_d, ro, rd, rot, rdt, z, bc, zs, za, sig, sig_t)
            for sub, img in (('render_com', person), ('render_head', head)):
                ref8 = O.to8b(img.numpy()).reshape(H, W, 3)
                ref_file = tmp_path / 'ref.jpg'
                Image.fromarray(ref8).save(str(ref_file))
                got = np.asarray(Image.open(tmp_path / sub / ('test_%06d.jpg' % i))).astype(np.int32)
                want = np.asarray(Image.open(ref_file)).astype(np.int32)
                # a 1e-6 error can flip a to8b truncation by one level in a few pixels; through JPEG that stays a few levels
                assert got.shape == (H, W, 3) and np.abs(got - want).max() <= 6 and (got != want).mean() < 0.05, (i, sub)
                seen.append(sub)
    assert len(seen) == 2 * n

"""CPU: the input formats (load_audface_data_split vs the REFERENCE loader's outputs on tests/golden/audface_tiny,
stored by oracle/make_golden_audface.py) and the output side (FrameWriter JPEG / MP4 files)."""
import os

import numpy as np
import pytest
import torch

import dfa_nerf_b200 as dfn

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

MODES = {   # must match oracle/make_golden_audface.py
    'test': dict(testskip=1, test_file='transforms_val_ba.json', aud_file='aud.pt', test_offset=0),
    'test_skip_off': dict(testskip=2, test_file='transforms_val_ba.json', aud_file='aud.pt', test_offset=3),
    'train_ba': dict(testskip=1, aud_file='aud.pt', use_ba=True),
    'train_skip': dict(testskip=2, aud_file='aud.pt', use_ba=True, no_com=True, all_speaker=True),
    'train_ori': dict(testskip=0, aud_file='aud.pt', use_ori=True),
}


@pytest.mark.parametrize('mode', sorted(MODES))
def test_loader_matches_reference(mode, monkeypatch):
    gold = np.load(os.path.join(GOLD, 'load_audface.npz'))
    monkeypatch.chdir(GOLD)                      # the golden image paths are relative to tests/golden
    d = dfn.load_audface_data_split('audface_tiny', **MODES[mode])
    keys = {k.split('/', 1)[1] for k in gold.files if k.startswith(mode + '/')}
    assert {('i_split0' if k == 'i_split' else k) for k in d} | ({'i_split1'} if 'i_split' in d else set()) == keys
    for k, v in d.items():
        if k == 'i_split':
            for j in range(2):
                g = gold['%s/i_split%d' % (mode, j)]
                assert v[j].dtype == g.dtype and np.array_equal(v[j], g)
            continue
        g = gold['%s/%s' % (mode, k)]
        if v is None:
            assert g.shape == () and str(g) == 'None', k
            continue
        v = np.asarray(v)
        assert v.shape == g.shape and v.dtype == g.dtype, (k, v.shape, g.shape, v.dtype, g.dtype)
        assert np.array_equal(v, g), k


def test_loader_edge_cases():
    base = os.path.join(GOLD, 'audface_tiny')
    d = dfn.load_audface_data_split(base, test_file='transforms_val_ba.json', aud_file='aud.pt')
    aud = torch.load(os.path.join(base, 'aud.pt')).numpy()
    exp = torch.load(os.path.join(base, 'face.pt'))['exp_o'].numpy()
    # val img_ids 9..13 run past both tables (11 and 9 rows): the last row is reused
    assert np.array_equal(d['auds'][-1], aud[-1]) and np.array_equal(d['auds'][0], aud[9])
    assert all(np.array_equal(e, exp[-1]) for e in d['exp'])
    assert d['hwfcxy'][:2] == [12, 10] and d['bc_img'].dtype == np.uint8 and d['bc_img'].shape == (12, 10, 3)
    pb = dfn.pose_body(base, use_ba=True)
    assert pb.shape == (4, 4) and pb.dtype == torch.float32 and torch.equal(pb, torch.eye(4) + torch.tensor([[0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, .5], [0, 0, 0, 0]]))
    with pytest.raises(dfn.DfnError):
        dfn.dataset_to_device(d, 'cpu')
    with pytest.raises(FileNotFoundError):
        dfn.load_audface_data_split(base, test_file='missing.json', aud_file='aud.pt')


def test_frame_writer_jpeg_and_video(tmp_path):
    from PIL import Image
    H, W, n = 48, 64, 6
    y, x = np.mgrid[0:H, 0:W]
    frames = [np.stack([(x * 3 + 10 * i) % 256, (y * 4) % 256, np.full_like(x, 40 * i)], -1).astype(np.uint8) for i in range(n)]
    w = dfn.FrameWriter(str(tmp_path / 'render_com'), workers=3, video=str(tmp_path / 'out.mp4'))
    for i in (3, 0, 5, 1, 4, 2):                           # any completion order; global indices name the files
        w.write(i, frames[i])
    paths = w.close()
    assert [os.path.basename(p) for p in paths] == ['test_%06d.jpg' % i for i in range(n)]
    for i, p in enumerate(paths):
        back = np.asarray(Image.open(p)).astype(np.int32)
        assert back.shape == (H, W, 3)
        assert np.abs(back - frames[i]).mean() < 6.0       # JPEG q75 of a smooth ramp
        ref = tmp_path / ('ref_%d.jpg' % i)
        Image.fromarray(frames[i]).save(str(ref))          # byte-identical to a plain Pillow save (what imageio.imwrite does)
        assert open(p, 'rb').read() == open(str(ref), 'rb').read()
    import cv2
    cap = cv2.VideoCapture(str(tmp_path / 'out.mp4'))
    cnt = 0
    while cap.read()[0]:
        cnt += 1
    assert cnt == n and abs(cap.get(cv2.CAP_PROP_FPS) - 25) < 1e-3
    with pytest.raises(ValueError):
        dfn.FrameWriter(str(tmp_path / 'x')).write(0, np.zeros((4, 4, 3), np.float32))

"""Build the tiny preprocessed-identity fixture tests/golden/audface_tiny/ and pin the data loader against the reference.

Run in the build container only (needs /root/reference):
    python oracle/make_golden_audface.py

Test infrastructure (like everything under oracle/).  The reference's loader (load_audface.py) imports `imageio`, which
this image does not have; imageio v2's `imread` of a JPEG is Pillow's decoder, so the stub below is
`np.asarray(PIL.Image.open(path))`.  The outputs of the REFERENCE loader on the fixture, in every mode the reference's
call site uses (MAIN:461-465), are stored in tests/golden/load_audface.npz.
"""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, 'tests', 'golden', 'audface_tiny')
REF = '/root/reference/NeRFs/DFANeRF'

MODES = {   # name -> kwargs of load_audface_data_split
    'test': dict(testskip=1, test_file='transforms_val_ba.json', aud_file='aud.pt', test_offset=0),
    'test_skip_off': dict(testskip=2, test_file='transforms_val_ba.json', aud_file='aud.pt', test_offset=3),
    'train_ba': dict(testskip=1, aud_file='aud.pt', use_ba=True),
    'train_skip': dict(testskip=2, aud_file='aud.pt', use_ba=True, no_com=True, all_speaker=True),
    'train_ori': dict(testskip=0, aud_file='aud.pt', use_ori=True),
}


def write_fixture():
    from PIL import Image
    os.makedirs(FIX, exist_ok=True)
    rng = np.random.default_rng(20260101)
    T = 11                                              # frames of audio; the expression file is shorter (clamping path)
    torch.save(torch.from_numpy(rng.standard_normal((T, 16, 29)).astype(np.float32)), os.path.join(FIX, 'aud.pt'))
    torch.save({'exp_o': torch.from_numpy(rng.standard_normal((T - 2, 64)).astype(np.float32))}, os.path.join(FIX, 'face.pt'))
    y, x = np.mgrid[0:12, 0:10]
    bc = np.stack([x * 25, y * 20, (x + y) * 11], -1).astype(np.uint8)
    Image.fromarray(bc).save(os.path.join(FIX, 'bc.jpg'), quality=92)
    np.save(os.path.join(FIX, 'speak_time.npy'), np.array([[0.0, 0.2], [0.27, 0.5]]))

    def frame(i, aud_id):
        a = 0.1 * i
        m = np.eye(4)
        m[:3, :3] = [[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]
        m[:3, 3] = [0.01 * i, -0.02 * i, 0.5 + 0.003 * i]
        return {'img_id': i, 'aud_id': aud_id, 'transform_matrix': m.tolist(), 'face_rect': [i, 2 * i, 5 + i, 6 + i]}

    for suffix in ('', '_ba'):
        k = 1.0 if suffix else 1.01
        for split, ids in (('train', range(0, 9)), ('val', range(9, 14))):      # val ids run past both feature files
            meta = {'focal_len': 1200.0 * k, 'cx': 5.0 * k, 'cy': 6.0, 'frames': [frame(i, min(i + 1, 12)) for i in ids]}
            with open(os.path.join(FIX, 'transforms_%s%s.json' % (split, suffix)), 'w') as fp:
                json.dump(meta, fp, indent=1)


def main():
    write_fixture()
    sys.path.insert(0, REF)
    imageio = types.ModuleType('imageio')
    from PIL import Image
    imageio.imread = lambda p: np.asarray(Image.open(p))
    sys.modules['imageio'] = imageio
    import load_audface as LOAD
    out = {}
    cwd = os.getcwd()
    os.chdir(os.path.dirname(FIX))       # the returned image PATHS contain basedir: keep them relative to tests/golden
    try:
        for name, kw in MODES.items():
            d = LOAD.load_audface_data_split('audface_tiny', **kw)
            for k, v in d.items():
                if v is None:
                    out['%s/%s' % (name, k)] = np.array('None')
                elif k == 'i_split':
                    out['%s/i_split0' % name], out['%s/i_split1' % name] = v
                else:
                    out['%s/%s' % (name, k)] = np.asarray(v)
            print('  ok  %-14s %s' % (name, sorted(d)))
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'load_audface.npz'), **out)


if __name__ == '__main__':
    main()

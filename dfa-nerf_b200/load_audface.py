"""Input formats of a preprocessed identity (reference load_audface.py:11-175, `LOAD`).

A data directory holds

* ``transforms_{train,val}[_ba].json`` (or the ``test_file``): ``focal_len``, ``cx``, ``cy`` and ``frames`` =
  ``[{img_id, aud_id, transform_matrix[4][4], face_rect[4]}]``;
* ``aud_file``  -- ``torch.save`` of the per-frame audio features (``[T,16,29]`` DeepSpeech windows or ``[T,512]``);
* ``exp_file``  -- ``torch.save`` of a dict whose ``'exp_o'`` is the per-frame expression code ``[T,64]``;
* ``bc.jpg``    -- the static background, which also fixes ``H, W``;
* ``speak_time.npy`` -- ``[k,2]`` speaking intervals in seconds (only without ``all_speaker``);
* ``head_imgs/ com_imgs/ ori_imgs/%06d.jpg`` -- training targets, referenced by PATH only (the reference reads them
  from disk per training step, MAIN:771-774; training is not built).

``load_audface_data_split`` has the reference's name, arguments and returned dict (same keys, dtypes and values: pinned
against the reference's loader on ``tests/golden/audface_tiny`` by ``oracle/make_golden_audface.py``).  It is host-side
table assembly -- no device work -- so it is plain numpy; ``dataset_to_device`` is the single upload MAIN:475-480 does,
and ``pose_body`` the fixed torso pose MAIN:453-459 reads.
"""
import json
import os

import numpy as np
import torch

_FPS = 30   # LOAD:151 (speak_time.npy is in seconds of a 30 fps clip)


def _read_json(path):
    with open(path) as fp:
        return json.load(fp)


def _read_tables(basedir, aud_file, exp_file, exp_offset=0):
    """The two per-frame feature tables (LOAD:20-21, 64-65)."""
    exp = torch.load(os.path.join(basedir, exp_file), map_location='cpu')['exp_o'].numpy()[exp_offset:]
    aud = torch.load(os.path.join(basedir, aud_file), map_location='cpu').cpu().numpy()
    return aud, exp


def _read_background(basedir):
    """bc.jpg as uint8 [H,W,3] RGB.  The reference calls imageio.imread (LOAD:36, 139), i.e. Pillow's decoder: same here."""
    from PIL import Image
    with Image.open(os.path.join(basedir, 'bc.jpg')) as im:
        return np.asarray(im)


def _gather_rows(table, ids):
    """table[min(id, T-1)] per frame -- ids past the end of a feature file reuse its last row (LOAD:25-30, 104-109)."""
    return table[np.minimum(np.asarray(ids, dtype=np.int64), table.shape[0] - 1)].astype(np.float32)


def _intrinsics(meta, bc_img):
    return [bc_img.shape[0], bc_img.shape[1], float(meta['focal_len']), float(meta['cx']), float(meta['cy'])]


def _frame_poses(frames):
    return np.array([f['transform_matrix'] for f in frames]).astype(np.float32)


def load_audface_data_split(basedir, testskip=1, test_file=None, aud_file=None, exp_file='face.pt',
                            no_com=False, all_speaker=False, use_ori=False, use_ba=False, test_offset=0):
    """LOAD:11-175.  With ``test_file`` (the render-only path, LOAD:14-47) the dict holds poses/auds/bc_img/hwfcxy/exp of
    every ``testskip``-th frame, BOTH tables indexed by ``img_id``; without it (LOAD:49-175) train+val are concatenated,
    audio is indexed by ``aud_id``, image PATHS are returned and ``i_split``/``speak_frames``/``sample_rects`` added."""
    if test_file:
        meta = _read_json(os.path.join(basedir, test_file))
        aud, exp = _read_tables(basedir, aud_file, exp_file, test_offset)
        frames = meta['frames'][::testskip]
        ids = [f['img_id'] for f in frames]
        bc_img = _read_background(basedir)
        return {'poses': _frame_poses(frames), 'auds': _gather_rows(aud, ids), 'bc_img': bc_img,
                'hwfcxy': _intrinsics(meta, bc_img), 'exp': _gather_rows(exp, ids)}

    suffix = '_ba' if use_ba else ''
    aud, exp = _read_tables(basedir, aud_file, exp_file)
    cols = {k: [] for k in ('imgs', 'imgs_com', 'imgs_ori', 'poses', 'auds', 'exp', 'sample_rects')}
    bounds = [0]
    for split in ('train', 'val'):
        meta = _read_json(os.path.join(basedir, 'transforms_%s%s.json' % (split, suffix)))
        step = 1 if (split == 'train' or testskip == 0) else testskip
        frames = meta['frames'][::step]
        img_ids = [f['img_id'] for f in frames]
        for key, sub in (('imgs', 'head_imgs'), ('imgs_com', 'com_imgs'), ('imgs_ori', 'ori_imgs')):
            cols[key].append(np.array([os.path.join(basedir, sub, '%06d.jpg' % i) for i in img_ids]))
        cols['poses'].append(_frame_poses(frames))
        cols['auds'].append(_gather_rows(aud, [f['aud_id'] for f in frames]))
        cols['exp'].append(_gather_rows(exp, img_ids))
        cols['sample_rects'].append(np.array([f['face_rect'] for f in frames], dtype=np.int32).reshape(len(frames), -1))
        bounds.append(bounds[-1] + len(frames))
    data = {k: np.concatenate(v, 0) for k, v in cols.items()}
    if no_com:
        data['imgs_com'] = None
    if not use_ori:
        data['imgs_ori'] = None

    n = data['auds'].shape[0]
    if all_speaker:
        speak = np.ones(n, dtype=np.int32)
    else:
        # frames strictly inside each [t0,t1] interval, one frame of margin at both ends (LOAD:148-154)
        speak = np.zeros(n, dtype=np.int32)
        for t0, t1 in np.load(os.path.join(basedir, 'speak_time.npy')):
            speak[np.arange(int(t0 * _FPS) + 1, int(t1 * _FPS) - 1)] = 1
    bc_img = _read_background(basedir)
    data.update(bc_img=bc_img, hwfcxy=_intrinsics(meta, bc_img), speak_frames=speak,     # intrinsics of the LAST split read
                i_split=[np.arange(bounds[i], bounds[i + 1]) for i in range(2)])
    return data


def pose_body(basedir, use_ba=False):
    """The fixed torso pose: frame 0 of the training transforms (MAIN:453-459), float32 [4,4]."""
    meta = _read_json(os.path.join(basedir, 'transforms_train%s.json' % ('_ba' if use_ba else '')))
    return torch.tensor(meta['frames'][0]['transform_matrix'], dtype=torch.float32)


def dataset_to_device(data, device, near=None, far=None):
    """MAIN:464-480: poses/auds/exp as float32 device tensors, the background scaled to [0,1]; one upload each."""
    if torch.device(device).type != 'cuda':
        from ._lib import DfnError
        raise DfnError('dfa_nerf_b200 has no CPU path: dataset_to_device needs a CUDA device')
    out = dict(data)
    for k in ('poses', 'auds', 'exp'):
        if out.get(k) is not None:
            out[k] = torch.from_numpy(np.ascontiguousarray(out[k])).to(device, torch.float32)
    out['bc_img'] = torch.from_numpy(np.array(data['bc_img'])).to(device).float() / 255.0    # copy: Pillow's array is read-only
    if near is not None:
        out['near'], out['far'] = near, far
    return out

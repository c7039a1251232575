"""The reference's `--render_person` job end to end (MAIN:590-733): a preprocessed identity directory in, JPEG frames out.

    data   = load_audface_data_split(datadir, testskip, test_file=..., aud_file=..., exp_file=...)     # LOAD:14-47
    paths  = render_person(decoder, data, pose_body(datadir, use_ba), z_shape, z_app, AudNet, ExpNet, out_dir, near, far)

One pre-pass encodes every frame's latents (encoders.encode_signal_sequence / encode_signal_torso_sequence: six launches
for the sequence instead of dozens per frame), the frame loop is sequence.render_sequence_head_torso (frames sharded
over ranks when torch.distributed is initialised), and finished frames are JPEG-encoded by frame_io.FrameWriter on host
threads while the next frames render.  Directory layout as the reference: ``<out_dir>/render_com/test_%06d.jpg`` (the
composite `person` image) and ``<out_dir>/render_head/test_%06d.jpg`` (MAIN:597-603, 717-722).
"""
import os

import torch

from . import _lib
from .encoders import encode_signal_sequence, encode_signal_torso_sequence
from .frame_io import FrameWriter
from .load_audface import dataset_to_device
from .sequence import render_sequence_head_torso


@torch.no_grad()
def render_person(decoder, data, pose_body, z_shape, z_app, AudNet, ExpNet, out_dir, near, far, AudAttNet=None, PoseAttNet=None,
                  N_samples=64, precision=_lib.PREC_BF16X3, device='cuda', video=None, workers=4, group=None):
    """data: the dict load_audface_data_split returns (numpy) or dataset_to_device's copy of it.  AudAttNet / PoseAttNet
    None = the `global_step < nosmo_iters` branch (MAIN:31, 81).  Returns this rank's written `render_com` paths."""
    if not torch.is_tensor(data['poses']):
        data = dataset_to_device(data, device)
    H, W, focal, cx, cy = data['hwfcxy']
    H, W = int(H), int(W)
    poses = data['poses']
    signals = encode_signal_sequence(data['auds'], data['exp'], AudNet, ExpNet, AudAttNet)
    signals_torso = encode_signal_torso_sequence(poses, PoseAttNet)
    com = FrameWriter(os.path.join(out_dir, 'render_com'), workers=workers, video=video)
    head = FrameWriter(os.path.join(out_dir, 'render_head'), workers=workers)

    def on_frame(i, planes):
        head.write(i, planes[0])
        com.write(i, planes[1])

    render_sequence_head_torso(decoder, H, W, focal, poses, pose_body.to(poses.device), data['bc_img'], z_shape, z_app, signals,
                               signals_torso, near, far, cx, cy, N_samples=N_samples, precision=precision, group=group,
                               gather=False, on_frame=on_frame, with_head=True, keep=False)
    head.close()
    return com.close()

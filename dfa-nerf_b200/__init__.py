"""dfa-nerf_b200: B200-native (sm_100a) volume-rendering hot path for DFA-NeRF.

Host-side mirror of the reference's call surface over the C ABI of
libdfn.so (include/dfn.h).  There is NO CPU path: every function needs CUDA tensors and the
compiled extension, and fails loudly without them.
"""
from ._lib import lib, DfnError, PREC_FP32, PREC_BF16, PREC_FP16, PREC_BF16X3, PREC_FP16X3M  # noqa: F401
from .functional import (get_rays, get_embedder, Embedder, decoder_transform_points, z_vals_uniform, make_points,  # noqa: F401
                         calc_volume_weights, composite_function, raw2outputs, sample_pdf, invert_cdf,
                         sort_merge, to8b, coarse_to_fine)
from .modules import NeRF, FaceNeRF  # noqa: F401
from .decoder import Decoder, DeformationField_ori, render_head_torso  # noqa: F401
from .render import render, render_rays, batchify_rays, run_network, RenderEngine  # noqa: F401
from .distributed import render_sharded, shard_range, gather_rgb, RayShardSink  # noqa: F401
from .encoders import (AudioNet, AudioNet_W2L, ExpressionEnc, AudioAttNet, encode_signal, encode_signal_torso,  # noqa: F401
                       encode_signal_sequence, encode_signal_torso_sequence, pose_to_euler_trans)
from .sequence import render_sequence, render_sequence_head_torso, shard_frames, FrameSink  # noqa: F401
from .load_audface import load_audface_data_split, dataset_to_device, pose_body  # noqa: F401
from .frame_io import FrameWriter, write_video  # noqa: F401
from .render_person import render_person  # noqa: F401
from .train import Trainer, select_coords  # noqa: F401

__all__ = ['get_rays', 'get_embedder', 'Embedder', 'decoder_transform_points', 'z_vals_uniform', 'make_points',
           'calc_volume_weights', 'composite_function', 'raw2outputs', 'sample_pdf', 'invert_cdf', 'sort_merge', 'coarse_to_fine',
           'NeRF', 'FaceNeRF', 'Decoder', 'DeformationField_ori', 'render_head_torso', 'render', 'render_rays', 'batchify_rays', 'run_network', 'RenderEngine',
           'render_sharded', 'shard_range', 'gather_rgb', 'RayShardSink', 'AudioNet', 'AudioNet_W2L', 'ExpressionEnc', 'AudioAttNet', 'encode_signal', 'encode_signal_torso', 'encode_signal_sequence', 'encode_signal_torso_sequence', 'pose_to_euler_trans', 'render_sequence', 'render_sequence_head_torso', 'shard_frames', 'FrameSink', 'to8b', 'load_audface_data_split', 'dataset_to_device', 'pose_body', 'FrameWriter', 'write_video', 'render_person', 'Trainer', 'select_coords', 'lib', 'DfnError', 'PREC_FP32', 'PREC_BF16', 'PREC_FP16', 'PREC_BF16X3', 'PREC_FP16X3M']

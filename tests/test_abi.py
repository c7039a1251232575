"""CPU: libdfn.so loads and exports every symbol include/dfn.h declares; host-side argument
checks run without a GPU (no compute calls)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'dfn.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(dfn_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported():
    from dfa_nerf_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), 'libdfn.so does not export %s' % n
    assert set(_lib.EXPORTS) == set(names), set(_lib.EXPORTS) ^ set(names)
    assert _lib.lib.dfn_abi_version() == 1


def test_no_cpu_path():
    import dfa_nerf_b200 as d
    with pytest.raises(d.DfnError):
        d.get_embedder(10)[0](torch.zeros(4, 3))
    with pytest.raises(d.DfnError):
        d.sample_pdf(torch.zeros(2, 63), torch.zeros(2, 62), 8, det=True)
    with pytest.raises(d.DfnError):
        d.FaceNeRF(D=8, W=256, input_ch=63, input_ch_views=27, dim_aud=64, use_viewdirs=True)(torch.zeros(3, 154))


def test_host_argument_checks():
    from dfa_nerf_b200 import _lib
    lib = _lib.lib
    assert lib.dfn_embed(0, None, 10, 0, None, None) == -1
    assert b'dfn_embed' in lib.dfn_last_error()
    assert lib.dfn_sample_pdf(4, 1000, None, None, 0, 8, None, 0, None, None, None) == -1
    desc = _lib.ModelDesc(7, 8, 256, 63, 27, 64, 4, 10, 4)
    h = ctypes.c_void_p()
    assert lib.dfn_model_create(ctypes.byref(desc), ctypes.byref(h)) == -1
    desc = _lib.ModelDesc(0, 8, 256, 63, 27, 64, 4, 10, 4)
    assert lib.dfn_model_create(ctypes.byref(desc), ctypes.byref(h)) == 0
    assert lib.dfn_model_num_tensors(h) == 2 * (8 + 3 + 3)
    assert lib.dfn_mlp_workspace_bytes(h, 1000) >= 2 * 1000 * 256 * 4
    # forward before load must fail with a state error, not touch the device
    assert lib.dfn_mlp_forward(h, 8, ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), 1 << 20, None) == -2
    lib.dfn_model_destroy(h)


def test_module_state_dict_keys_match_reference():
    """Parameter names equal the reference's (HELP:257-273, HELP:354-370), so checkpoints load unchanged."""
    import dfa_nerf_b200 as d
    from oracle import synth
    m = d.FaceNeRF(D=8, W=256, input_ch=63, input_ch_views=27, dim_aud=64, output_ch=4, skips=[4], use_viewdirs=True)
    sd = synth.facenerf_state_dict(0)
    assert set(m.state_dict().keys()) == set(sd.keys())
    m.load_state_dict(sd)
    n = d.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=4, skips=[4], use_viewdirs=True)
    sdn = synth.nerf_state_dict(0)
    assert set(n.state_dict().keys()) == set(sdn.keys())
    n.load_state_dict(sdn)
    assert sum(p.numel() for p in m.parameters()) == 661636
    assert sum(p.numel() for p in n.parameters()) == 595844


def test_shard_range_covers_all_rays():
    from dfa_nerf_b200.distributed import shard_range
    for n in (202500, 4096, 7, 1):
        for g in (1, 2, 4, 8):
            seen = []
            for r in range(g):
                b, e, per = shard_range(n, r, g)
                assert e - b <= per
                seen += list(range(b, e)) if n < 10000 else [b, e]
            if n < 10000:
                assert seen == list(range(n))
    assert shard_range(202500, 7, 8) == (177191, 202500, 25313)

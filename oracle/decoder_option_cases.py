"""Test infrastructure: the Decoder configurations of tests/golden/decoder_options.npz (written by
oracle/make_golden_decoder_options.py from the reference's own decoder.Decoder, DEC:166-349)."""

# name -> (constructor keywords, [(call name, head_or_torso, has signal, has expression, has ray_d)])
CASES = {
    'expr': (dict(hidden_size=64, z_dim=32, dim_signal=24, dim_exp=40, dim_et_embed=18, n_blocks=6, skips=[2, 4],
                  use_deformation_field=True, use_expression=True, use_wav2lip=True, dim_w2lfeature=8),
             [('head_expr', 'head', True, True, True), ('head_noexpr', 'head', True, False, True),
              ('listener_expr', 'head', False, True, True), ('listener', 'head', False, False, True),
              ('torso', 'torso', True, False, True)]),
    'plain': (dict(hidden_size=96, z_dim=16, dim_signal=10, dim_et_embed=7, n_blocks=5, skips=[1], n_freq_posenc=6,
                   n_freq_posenc_views=2, rgb_out_dim=5, final_sigmoid_activation=False, use_deformation_field=False),
              [('head_noview', 'head', True, False, False), ('head', 'head', True, False, True),
               ('torso_nodeform', 'torso', True, False, True), ('listener_noview', 'head', False, False, False)]),
}


def oracle_kwargs(kw, has_ex):
    """Keywords of oracle.nerf_oracle.decoder_forward for a constructor keyword set."""
    return dict(n_freq=kw.get('n_freq_posenc', 10), n_freq_views=kw.get('n_freq_posenc_views', 4), skips=tuple(kw['skips']),
                n_blocks=kw['n_blocks'], use_deformation_field=kw['use_deformation_field'],
                final_sigmoid=kw.get('final_sigmoid_activation', True))

import torch, sys
sys.path.insert(0, ".")
import dfa_nerf_b200 as dfn
import synth
dev = torch.device("cuda", 0)
net = dfn.FaceNeRF(D=8, W=256, input_ch=63, input_ch_views=27, dim_aud=64, output_ch=4, skips=[4], use_viewdirs=True)
net.load_state_dict(synth.facenerf_state_dict(1)); net = net.to(dev)
R, S = int(sys.argv[1]), 192
fr = synth.frame_inputs(H=450, W=450, seed=0)
ro, rd, vd = dfn.get_rays(450, 450, fr["focal"], fr["c2w"], fr["cx"], fr["cy"], device=dev, return_viewdirs=True)
ro, rd, vd = [t.reshape(-1, 3)[:R].contiguous() for t in (ro, rd, vd)]
z, _ = torch.sort(torch.rand(R, S, device=dev) * 0.6 + 0.4, -1)
eng = dfn.RenderEngine(net, None, S, 0, precision=dfn.PREC_BF16)
raw = eng.query_points(net, ro, rd, vd, z, fr["aud"].to(dev))
torch.cuda.synchronize()
print("done", R, raw.abs().mean().item())

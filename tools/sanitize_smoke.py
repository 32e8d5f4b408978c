import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scade_b200 import synthetic as syn, render as R_, nerf_helpers as NH
from tests.golden.generate_goldens import net_pair
dev = torch.device("cuda:0")
pc, pf = net_pair(8, 256)
def mk(p):
    net = NH.NeRF(D=8, W=256, input_ch=57, input_ch_views=3, output_ch=5, skips=[4], use_viewdirs=True, precision="tc_f16")
    net.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
    return net.to(dev)
bb_center, bb_scale = syn.bounding_box()
qf = R_.NetworkQuery(NH.get_embedder(9, 0)[0], NH.get_embedder(0, 0)[0], bb_center, bb_scale, precision="tc_f16")
kw = dict(network_fn=mk(pc), network_query_fn=qf, N_samples=24, embedded_cam=torch.tensor((), device=dev), retraw=False, perturb=0.0,
          N_importance=40, network_fine=mk(pf), raw_noise_std=0.0)
rb = torch.from_numpy(syn.make_ray_batch(77, seed=3)).to(dev)
with torch.no_grad():
    out = R_.render_rays(rb, True, **kw)
torch.cuda.synchronize()
print("eval ok", float(out["rgb_map"].sum()))
kw["perturb"] = 1.0
out = R_.render_rays(rb, True, **kw)
(out["rgb_map"].sum() + out["rgb0"].sum() + out["pred_hyp"].sum()).backward()
torch.cuda.synchronize()
print("train ok", float(kw["network_fine"].pts_linears[3].weight.grad.abs().sum()))

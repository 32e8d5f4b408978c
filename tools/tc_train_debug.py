"""Stage-by-stage check of the tensor-core training path against the oracle (GPU box only):
forward stash (activations, masks, alpha), dgrad stash (dZ of every layer), final gradients.

    python tools/tc_train_debug.py [P]
"""
import os
import sys
from ctypes import byref, c_void_p

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import scade_oracle as O                      # noqa: E402
from scade_b200 import _lib, functional as F_, nerf_helpers as NH, synthetic as syn   # noqa: E402


from tests.util import stash_unswizzle as unswizzle          # noqa: E402


def decode_mask(words):
    """sign_mask32 words [..., n] uint32 -> bool [..., 32 n] (True = sign bit set = inactive)."""
    k = np.arange(32)
    elem = 4 * (k & 7) + (k >> 3)
    out = np.zeros(words.shape[:-1] + (words.shape[-1], 32), bool)
    for kk in range(32):
        out[..., elem[kk]] = (words >> np.uint32(31 - kk)) & 1
    return out.reshape(words.shape[:-1] + (words.shape[-1] * 32,))


def rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 700
    dev = torch.device("cuda:0")
    D, W = 8, 256
    params = syn.make_nerf_params(seed=12, D=D, W=W, bias_scale=0.1, alpha_bias=0.3)
    net = NH.NeRF(D=D, W=W, input_ch=57, input_ch_views=3, output_ch=5, skips=[4], use_viewdirs=True, precision="tc_f16")
    net.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    net = net.to(dev)
    rng = np.random.default_rng(13)
    x = rng.uniform(-1, 1, (P, 60)).astype(np.float32)
    d_out = (rng.standard_normal((P, 4)) * 1e-4).astype(np.float32)
    h = net.handle()
    L = _lib.load()
    prec = _lib.PREC_TC_F16
    ws = torch.zeros(h.workspace_bytes(P, prec, 1), dtype=torch.uint8, device=dev)
    out = torch.empty((P, 4), dtype=torch.float32, device=dev)
    xs = torch.from_numpy(x).to(dev)
    cnet = h.struct(prec)
    _lib.check(L.scade_mlp_forward_embedded(byref(cnet), prec, _lib.ptr(xs), P, _lib.ptr(out), _lib.ptr(ws), ws.numel(), 1,
                                            _lib.stream_ptr()), "fwd")
    torch.cuda.synchronize()
    lay = F_.stash_layout(h, P)
    T = lay["T"]
    print("layout T =", T, "total MB =", lay["total"] / 2 ** 20)
    ref_out, acts = O.nerf_forward(params, x, return_acts=True)
    print("out   rel err vs fp32 oracle:", rel(out.cpu().numpy(), ref_out))
    buf = ws.cpu().numpy()
    emb = unswizzle(buf, lay["emb"], T, 1)[:P]
    print("emb   rel err:", rel(emb[:, :60], x), " ones:", emb[:, 60].min(), emb[:, 61].max())
    for l in range(D):
        hh = unswizzle(buf, lay["h"][l], T, 4)[:P]
        print(f"h[{l}]  rel err:", rel(hh, np.maximum(acts["pre"][l], 0)))
        mw = buf[lay["maskh"][l]:lay["maskh"][l] + T * 128 * 32].view(np.uint32).reshape(T, 2, 128, 4).transpose(0, 2, 1, 3).reshape(T * 128, 8)
        inact = decode_mask(mw)[:P]
        refm = acts["pre"][l] < 0
        print(f"mask[{l}] mismatches:", int((inact != refm).sum()), "of", refm.size,
              " (|pre| at mismatches max", float(np.abs(acts['pre'][l][inact != refm]).max()) if (inact != refm).any() else 0.0, ")")
    feat = unswizzle(buf, lay["feat"], T, 4)[:P]
    print("feat  rel err:", rel(feat, acts["hv_in"][:, :W]))
    hv = unswizzle(buf, lay["hv"], T, 2)[:P]
    print("hv    rel err:", rel(hv, acts["hv"]))
    al = buf[lay["alpha"]:lay["alpha"] + T * 128 * 4].view(np.float32)[:P]
    print("alpha rel err:", rel(al, acts["alpha"][:, 0]))
    mv = decode_mask(buf[lay["maskv"]:lay["maskv"] + T * 128 * 16].view(np.uint32).reshape(T * 128, 4))[:P]
    print("maskv mismatches:", int((mv != (acts["zv"] < 0)).sum()))

    # ---- backward ----
    grads = [torch.zeros_like(p) for p in h.params]
    arr = (c_void_p * len(grads))(*[g.data_ptr() for g in grads])
    ds = torch.from_numpy(d_out).to(dev)
    _lib.check(L.scade_mlp_backward(byref(cnet), prec, _lib.ptr(ds), P, arr, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "bwd")
    torch.cuda.synchronize()
    buf = ws.cpu().numpy()
    maxbits = int(buf[lay["gs"]:lay["gs"] + 4].view(np.uint32)[0])
    scale = 2.0 ** (5 - ((maxbits >> 23) - 127))
    print("scale", scale)
    # teacher-forced reference (fp64 arithmetic on the stashed fp16 activations and masks): isolates the backward kernels
    # from the sign flips the fp16 forward causes near z = 0
    p64 = {k: v.astype(np.float64) for k, v in params.items()}
    H = [unswizzle(buf, lay["h"][l], T, 4)[:P].astype(np.float64) for l in range(D)]
    inact = [decode_mask(buf[lay["maskh"][l]:lay["maskh"][l] + T * 128 * 32].view(np.uint32).reshape(T, 2, 128, 4).transpose(0, 2, 1, 3).reshape(T * 128, 8))[:P] for l in range(D)]
    feat64, hv64, emb64 = feat.astype(np.float64), hv.astype(np.float64), emb.astype(np.float64)
    d_rgb, d_sigma = d_out[:, :3].astype(np.float64), d_out[:, 3:4].astype(np.float64)
    al64 = al.astype(np.float64)[:, None]
    d_alpha = d_sigma * np.where(al64 * 10 > 20, 1.0, 1.0 / (1.0 + np.exp(-al64 * 10.0)))
    tf = {}
    d_zv = (d_rgb @ p64["rgb_linear.weight"]) * (~mv)
    print("dzv   rel err (teacher-forced):", rel(unswizzle(buf, lay["dzv"], T, 2)[:P] / scale, d_zv))
    tf["rgb_linear.weight"] = d_rgb.T @ hv64
    tf["rgb_linear.bias"] = d_rgb.sum(0)
    tf["views_linears.0.weight"] = d_zv.T @ np.concatenate([feat64, emb64[:, 57:60]], -1)
    tf["views_linears.0.bias"] = d_zv.sum(0)
    d_feat = (d_zv @ p64["views_linears.0.weight"])[:, :W]
    print("dzf   rel err (teacher-forced):", rel(unswizzle(buf, lay["dzf"], T, 4)[:P] / scale, d_feat))
    tf["feature_linear.weight"] = d_feat.T @ H[D - 1]
    tf["feature_linear.bias"] = d_feat.sum(0)
    tf["alpha_linear.weight"] = d_alpha.T @ H[D - 1]
    tf["alpha_linear.bias"] = d_alpha.sum(0)
    d_h = d_feat @ p64["feature_linear.weight"] + d_alpha @ p64["alpha_linear.weight"]
    for i in reversed(range(D)):
        d_z = d_h * (~inact[i])
        print(f"dz[{i}] rel err (teacher-forced):", rel(unswizzle(buf, lay["dz"][i], T, 4)[:P] / scale, d_z))
        xin = emb64[:, :57] if i == 0 else (np.concatenate([emb64[:, :57], H[i - 1]], -1) if i - 1 == 4 else H[i - 1])
        tf[f"pts_linears.{i}.weight"] = d_z.T @ xin
        tf[f"pts_linears.{i}.bias"] = d_z.sum(0)
        if i > 0:
            d_h = d_z @ p64[f"pts_linears.{i}.weight"]
            if i - 1 == 4:
                d_h = d_h[:, 57:]
    ref = O.nerf_backward(params, x, d_out, dtype=np.float64)
    names = [n for n, _ in net.named_parameters()]
    order = []
    for i in range(D):
        order += [f"pts_linears.{i}.weight", f"pts_linears.{i}.bias"]
    order += ["views_linears.0.weight", "views_linears.0.bias", "feature_linear.weight", "feature_linear.bias",
              "alpha_linear.weight", "alpha_linear.bias", "rgb_linear.weight", "rgb_linear.bias"]
    worst = 0.0
    for name, g in zip(order, grads):
        r = ref[name]
        gn = g.cpu().numpy().astype(np.float64)
        err, err_tf = rel(gn, r), rel(gn, tf[name].reshape(gn.shape))
        worst = max(worst, err_tf)
        cos = float((gn * r).sum() / (np.linalg.norm(gn) * np.linalg.norm(r) + 1e-30))
        print(f"grad {name:28s} vs teacher-forced {err_tf:.3e}   vs fp64 oracle {err:.3e} (cos {cos:.5f})  |ref|max {np.abs(r).max():.3e}")
    print("WORST grad rel err vs teacher-forced reference:", worst)


if __name__ == "__main__":
    main()

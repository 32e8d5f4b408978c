"""Edge cases of the hot path on the GPU, through the reference-named wrappers: ragged and tiny ray counts against the
tile sizes of the tensor-core kernel (128 / 256 / 512 points), empty densities (the reference's NaN disparity, RS:559, and
the `denom < 1e-5 -> 1` rule of sample_pdf, H:378-379), all-zero masks and thresholds in the space-carving loss (H:108-113),
sample counts that are not multiples of the warp width, and BASELINE config 1 (64 rays x 64 samples, raw2outputs only)."""
import numpy as np
import pytest
import torch

from oracle import scade_oracle as O
from scade_b200 import synthetic as syn
from tests.test_gpu_parity import T, close, dev, make_render_kwargs, npy  # noqa: F401  (dev is a fixture)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["tc_f16", "fp32"])
def test_ragged_ray_counts_are_ray_independent(dev, precision):
    """Rays are independent units: any sub-batch -- 1 ray, a count straddling the 128-point row tile, the
    256-point CTA step and the 512-point cluster step -- renders to exactly the values it has inside a larger batch."""
    from scade_b200 import render as R_
    Nc, Nf = 24, 40                                       # 24 and 64 samples per ray: P is rarely a multiple of 128
    kwargs, _ = make_render_kwargs(8, 256, dev, precision, 0.0, Nc, Nf)
    rb = T(syn.make_ray_batch(301, seed=7), dev)
    with torch.no_grad():
        full = R_.render_rays(rb, True, **kwargs)
        for n in (1, 2, 5, 11, 21, 22, 43, 129, 300):
            sub = R_.render_rays(rb[:n], True, **kwargs)
            for k in ("rgb_map", "depth_map", "acc_map", "z_vals", "weights", "pred_hyp", "raw", "rgb0", "z_std"):
                np.testing.assert_array_equal(npy(sub[k]), npy(full[k])[:n], err_msg=f"{k} n={n}")
    assert np.isfinite(npy(full["rgb_map"])).all()


def test_config1_raw2outputs_only(dev):
    """BASELINE config 1: 64 rays x 64 samples, raw = randn, z = linspace(0.1, 5, 64); compositing only."""
    from scade_b200 import render as R_
    rng = np.random.default_rng(0)
    raw = rng.standard_normal((64, 64, 4)).astype(np.float32)
    z = np.broadcast_to(np.linspace(0.1, 5.0, 64, dtype=np.float32), (64, 64)).copy()
    rays_d = syn.make_ray_batch(64, seed=3)[:, 3:6]
    ref = O.raw2outputs(raw, z, rays_d)
    out = R_.raw2outputs(T(raw, dev), T(z, dev), T(rays_d, dev))
    for a, b, name in zip(out, ref, ("rgb_map", "disp_map", "acc_map", "weights", "depth_map")):
        close(npy(a), b, rtol=2e-5, atol=2e-6)


def test_empty_density_gives_reference_nan_disparity_and_uniform_resampling(dev):
    """sigma = -inf side of the ReLU (RS:512): weights are exactly 0, acc = 0, depth = 0 and disp = 1/max(1e-10, 0/0) = NaN as in
    the reference (torch.max propagates NaN, RS:559).  sample_pdf on all-zero weights is the uniform pdf over
    the bins (w + 1e-5, H:339) and must agree with the oracle, including the `denom < 1e-5 -> 1` branch."""
    from scade_b200 import nerf_helpers as NH
    from scade_b200 import render as R_
    N, S = 37, 50                                        # S is not a multiple of 32
    raw = np.zeros((N, S, 4), np.float32)
    raw[..., 3] = -5.0
    raw[..., :3] = np.random.default_rng(1).standard_normal((N, S, 3)).astype(np.float32)
    z = np.sort(np.random.default_rng(2).uniform(0.1, 5.0, (N, S)).astype(np.float32), -1)
    rays_d = syn.make_ray_batch(N, seed=4)[:, 3:6]
    rgb, disp, acc, w, depth = [npy(t) for t in R_.raw2outputs(T(raw, dev), T(z, dev), T(rays_d, dev))]
    assert (w == 0).all() and (acc == 0).all() and (depth == 0).all() and (rgb == 0).all()
    assert np.isnan(disp).all()
    o = O.raw2outputs(raw, z, rays_d)
    assert np.isnan(o[1]).all()
    # resampling from empty weights
    bins = 0.5 * (z[:, 1:] + z[:, :-1])
    wmid = w[:, 1:-1]
    u = np.random.default_rng(3).uniform(0, 1, (N, 33)).astype(np.float32)
    ref, _ = O.sample_pdf(bins, wmid, 33, u=u)
    got, u_used = NH.sample_pdf_return_u(T(bins, dev), T(wmid, dev), 33, load_u=T(u, dev))
    close(npy(got), ref, rtol=1e-5, atol=1e-5)
    np.testing.assert_array_equal(npy(u_used), u)
    # one bin carries all the mass: every other cdf step is below 1e-5 (denom -> 1, H:378-379)
    spike = np.zeros_like(wmid)
    spike[:, 7] = 1.0e3
    ref, _ = O.sample_pdf(bins, spike, 33, u=u)
    got = NH.sample_pdf_return_u(T(bins, dev), T(spike, dev), 33, load_u=T(u, dev))[0]
    close(npy(got), ref, rtol=1e-5, atol=1e-5)
    # det=True uses linspace including both endpoints (H:347): first sample = first bin edge, last = last edge
    got = npy(NH.sample_pdf(T(bins, dev), T(wmid, dev), 33, det=True))
    close(got[:, 0], bins[:, 0], rtol=1e-6, atol=1e-6)
    close(got[:, -1], bins[:, -1], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("is_joint", [False, True])
def test_space_carving_masks_and_threshold(dev, is_joint):
    """H:108-113: an all-zero mask zeroes every distance (loss 0, zero gradients); a threshold above every distance does the
    same; a mixed mask and a mid threshold agree with the oracle.  K = 1 and a ragged P exercise the reduction tails."""
    from scade_b200 import nerf_helpers as NH
    rng = np.random.default_rng(5)
    for K, N, P in ((1, 9, 5), (20, 70, 33), (7, 130, 128)):
        pred = rng.uniform(0.1, 5.0, (N, P)).astype(np.float32)
        hyp = rng.uniform(0.1, 5.0, (K, N, 1)).astype(np.float32)
        for mask, thr in ((np.zeros(N, np.float32), 0.0), (None, 100.0), ((rng.uniform(size=N) > 0.5).astype(np.float32), 0.0),
                          (None, 0.7)):
            p_t = T(pred, dev).requires_grad_(True)
            h_t = T(hyp, dev).requires_grad_(True)
            loss = NH.compute_space_carving_loss(p_t, h_t, is_joint=is_joint, mask=None if mask is None else T(mask, dev),
                                                 threshold=thr)
            ref = O.space_carving_loss(pred, hyp, is_joint=is_joint, mask=mask, threshold=thr)
            close(float(loss.detach()), float(ref), rtol=2e-5, atol=1e-7)
            loss.backward()
            d_pred, d_hyp = O.space_carving_loss_bwd(pred, hyp, is_joint=is_joint, mask=mask, threshold=thr)[:2]
            if is_joint and thr == 0.7:
                continue        # arg-min over K of near-equal means may differ legitimately; the loss value is what is pinned
            close(npy(p_t.grad), d_pred, rtol=2e-5, atol=1e-8)
            close(npy(h_t.grad), np.asarray(d_hyp).reshape(hyp.shape), rtol=2e-4, atol=1e-7)
            if (mask is not None and not mask.any()) or thr == 100.0:
                assert float(loss) == 0.0 and not npy(p_t.grad).any() and not npy(h_t.grad).any()


def test_cpu_input_and_bad_shapes_fail_loudly(dev):
    """No CPU fallback and no silent reshaping: CPU tensors and N_importance == 0 raise."""
    from scade_b200 import render as R_
    from scade_b200._lib import ScadeError
    kwargs, _ = make_render_kwargs(8, 256, dev, "tc_f16", 0.0, 16, 16)
    rb = torch.from_numpy(syn.make_ray_batch(4, seed=1))
    with pytest.raises(ScadeError):
        R_.render_rays(rb, True, **kwargs)
    kwargs["N_importance"] = 0
    with pytest.raises(NotImplementedError):
        R_.render_rays(rb.to(dev), True, **kwargs)


def test_graphed_render_rays_equals_eager(dev):
    """GraphedRenderRays replays render_rays as one CUDA graph: same values as the eager call, for device and pinned-host inputs."""
    from scade_b200 import render as R_
    kwargs, _ = make_render_kwargs(8, 256, dev, "tc_f16", 0.0, 32, 48)
    kwargs["retraw"] = False
    g = R_.GraphedRenderRays(200, **kwargs)
    for seed in (11, 12):
        rb = syn.make_ray_batch(200, seed=seed)
        with torch.no_grad():
            ref = R_.render_rays(T(rb, dev), True, **kwargs)
        out = g(torch.from_numpy(rb).pin_memory() if seed == 12 else T(rb, dev))
        torch.cuda.synchronize()
        for k in ref:
            np.testing.assert_array_equal(npy(out[k]), npy(ref[k]), err_msg=k)
    with pytest.raises(ValueError):
        g(T(syn.make_ray_batch(10, seed=1), dev))
    # host-I/O form: the graph contains the H2D copy of rays_host and the D2H copies of the requested maps
    gh = R_.GraphedRenderRays(200, host_outputs=("rgb_map", "depth_map"), **kwargs)
    rb = syn.make_ray_batch(200, seed=13)
    gh(torch.from_numpy(rb))
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = R_.render_rays(T(rb, dev), True, **kwargs)
    np.testing.assert_array_equal(gh.out_host["rgb_map"].numpy(), npy(ref["rgb_map"]))
    np.testing.assert_array_equal(gh.out_host["depth_map"].numpy(), npy(ref["depth_map"]))


def test_sharded_image_render_with_graph_cache(dev):
    """render_image_sharded (single rank here): chunks replayed through cached CUDA graphs give the same maps as the eager
    batchify_rays path, including a ragged last chunk, and the cache is reused across frames."""
    from scade_b200.dist import render_image_sharded
    kwargs, _ = make_render_kwargs(8, 256, dev, "tc_f16", 0.0, 16, 24)
    kwargs["retraw"] = False
    cache = {}
    for pose in syn.spiral_poses(6)[1:3]:
        c2w = torch.from_numpy(pose)
        a = render_image_sharded(30, 41, syn.CAM_INTRINSIC, c2w, 0.1, 5.0, kwargs, chunk=500, graph_cache=cache)
        b = render_image_sharded(30, 41, syn.CAM_INTRINSIC, c2w, 0.1, 5.0, kwargs, chunk=500)
        for k in b:
            np.testing.assert_array_equal(npy(a[k]), npy(b[k]), err_msg=k)
    assert list(cache) == [500]


def test_cpp_host_renders_through_the_c_abi(dev):
    """The torch-free C++ host (examples/render_cabi.cpp) renders a small frame through libscade_b200.so and reports finite,
    in-range maps -- the C ABI is usable without Python."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "examples", "render_cabi")
    if not os.path.exists(exe):
        r = subprocess.run(["sh", os.path.join(root, "examples", "build.sh")], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([exe, "96", "128"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK" in r.stdout and "kernel launches/frame" in r.stdout, r.stdout


def test_pipelined_render_rays_matches_single_calls(dev):
    """PipelinedRenderRays (two graph slots on two streams, copies overlapping compute): every submission returns exactly what
    render_rays returns for that batch, in order, also when slots are reused."""
    import numpy as np
    import torch
    from scade_b200 import nerf_helpers as NH, render as R_, synthetic as syn
    from tests.golden.generate_goldens import net_pair
    pc, pf = net_pair(8, 256)
    bb_center, bb_scale = syn.bounding_box()

    def mk(p):
        net = NH.NeRF(D=8, W=256, input_ch=57, input_ch_views=3, output_ch=5, skips=[4], use_viewdirs=True, precision="tc_f16")
        net.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
        return net.to(dev).requires_grad_(False)
    qf = R_.NetworkQuery(NH.get_embedder(9, 0)[0], NH.get_embedder(0, 0)[0], bb_center, bb_scale, precision="tc_f16")
    kw = dict(network_fn=mk(pc), network_query_fn=qf, N_samples=64, embedded_cam=torch.tensor((), device=dev), perturb=0.0,
              N_importance=128, network_fine=mk(pf), raw_noise_std=0.0)
    n = 300
    batches = [torch.from_numpy(syn.make_ray_batch(n, seed=200 + i)).pin_memory() for i in range(5)]
    pipe = R_.PipelinedRenderRays(n, depth=2, host_outputs=("rgb_map", "depth_map"), **kw)
    tickets, got = [], []
    for i, b in enumerate(batches):
        tickets.append(pipe.submit(b))
        if i >= 1:                                        # consume with one submission in flight
            out = pipe.result(tickets[i - 1])
            got.append({k: v.clone() for k, v in out.items()})
    got.append({k: v.clone() for k, v in pipe.result(tickets[-1]).items()})
    with torch.no_grad():
        for b, g in zip(batches, got):
            ref = R_.render_rays(b.to(dev), True, **kw)
            np.testing.assert_array_equal(g["rgb_map"].numpy(), ref["rgb_map"].cpu().numpy())
            np.testing.assert_array_equal(g["depth_map"].numpy(), ref["depth_map"].cpu().numpy())
    with pytest.raises(ValueError):
        pipe.result(tickets[0])                           # that slot has been reused since

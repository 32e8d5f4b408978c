"""Golden fixtures for the SURVEY §8(f) rows (training sampler, video post-processing), made by EXECUTING THE UNMODIFIED
REFERENCE on CPU in the build container:

    python tests/golden/generate_goldens_f.py

Writes tests/golden/train_batch.npz and tests/golden/video_frame.npz.  Inputs are rebuilt in the tests from
scade_b200.synthetic (seeded), only the reference's outputs are stored.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden.generate_goldens import import_reference, t  # noqa: E402
from scade_b200 import synthetic as syn  # noqa: E402

TRAIN_SCENE = dict(n_img=2, H=48, W=64, K=3, n_u=5, seed=40)
N_RAND, IMG_I, SEED = 300, 1, 123
FRAME = dict(H=12, W=20, S=24, seed=41, depth_scale=5.0)


def frame_inputs():
    rng = np.random.default_rng(FRAME["seed"])
    H, W, S = FRAME["H"], FRAME["W"], FRAME["S"]
    rgb = rng.uniform(-0.1, 1.1, (H, W, 3)).astype(np.float32)
    z = np.sort(rng.uniform(0.1, 5.0, (H, W, S)), -1).astype(np.float32)
    w = rng.random((H, W, S)).astype(np.float32)
    w = (w / w.sum(-1, keepdims=True) * rng.uniform(0.2, 1.0, (H, W, 1))).astype(np.float32)
    depth = (w * z).sum(-1).astype(np.float32)
    depth[0, :3] = [-0.5, 0.0, 7.0]                     # exercise both clips of to8b / to16b
    return rgb, depth, z, w


def main():
    R, H_ = import_reference()
    sc = syn.make_train_scene(**TRAIN_SCENE)
    H, W = TRAIN_SCENE["H"], TRAIN_SCENE["W"]
    args = types.SimpleNamespace(N_rand=N_RAND, mask_corners=True)
    np.random.seed(SEED)
    out = R.get_ray_batch_from_one_image_hypothesis_idx(
        H, W, IMG_I, t(sc["images"]), t(sc["depths"]), t(sc["valid_depths"]), t(sc["poses"]), t(sc["intrinsics"]),
        t(sc["all_hypothesis"]), args, None, t(sc["cached_u"]))
    batch_rays, target_s, target_d, target_vd, img_i, target_h, mask, cu = out
    np.random.seed(SEED)
    sel = np.random.choice(H * W, size=[N_RAND], replace=False)
    np.savez_compressed(os.path.join(HERE, "train_batch.npz"), select_inds=sel, batch_rays=batch_rays.numpy(),
                        target_s=target_s.numpy(), target_d=target_d.numpy(), target_vd=target_vd.numpy(),
                        target_h=target_h.numpy(), mask=mask.numpy(), cached_u=cu.numpy())
    # ---- video frame pieces, exactly the expressions of render_video (RS:252-259) and write_images_with_metrics (RS:403) ----
    rgb, depth, z, w = frame_inputs()
    extras = {"depth_map": t(depth), "z_vals": t(z), "weights": t(w)}
    rgb8 = H_.to8b(t(rgb).cpu().numpy())
    d8 = H_.to8b((extras["depth_map"] / FRAME["depth_scale"]).cpu().numpy())
    depth_var = ((extras["z_vals"] - extras["depth_map"].unsqueeze(-1)).pow(2) * extras["weights"]).sum(-1)
    depth_std = depth_var.clamp(0., 1.).sqrt()
    s8 = H_.to8b(depth_std.cpu().numpy())
    d16 = H_.to16b(depth)
    np.savez_compressed(os.path.join(HERE, "video_frame.npz"), rgb8=rgb8, depth8=d8, std8=s8, depth_std=depth_std.numpy(),
                        depth16=d16)
    print("wrote train_batch.npz, video_frame.npz")


if __name__ == "__main__":
    main()

"""CPU-side checks of the C-ABI boundary: the library builds/loads and exports exactly the header's
symbols; argument validation returns error codes (no compute needs a GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from scade_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "scade_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(scade_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert sorted(_lib.exported_symbols()) == header_symbols()


def test_library_loads_and_exports_every_symbol():
    lib = _lib.load()
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert lib.scade_version() == 103


def test_argument_validation_without_gpu():
    lib = _lib.load()
    bad = _lib.NetDesc(8, 255, 9, 0, 4)      # odd width
    assert lib.scade_mlp_workspace_bytes(ctypes.byref(bad), 1024, 0, 0) == 0
    good = _lib.NetDesc(8, 256, 9, 0, 4)
    assert lib.scade_mlp_workspace_bytes(ctypes.byref(good), 1024, 0, 0) > 1024 * 256 * 4
    # 74 forward weight stages (2 + 6 x 8 + 10 + 8 + 6) + fp32 tail (heads + per-layer bias rows, 17 KB) + 68 dgrad (W^T) stages
    assert lib.scade_mlp_packed_bytes(ctypes.byref(good)) == 74 * 16384 + 17408 + 68 * 16384
    # training stash of the tensor-core path: 13 + 8 D chunks of 16 KB per 128-point tile + masks + alpha
    lay = (ctypes.c_int64 * 64)()
    assert lib.scade_mlp_tc_stash_layout(ctypes.byref(good), 1000, lay, 64) == 35
    assert lay[0] == 8 and lay[1] == 8 and lay[2] == lib.scade_mlp_workspace_bytes(ctypes.byref(good), 1000, 1, 1)
    assert lay[2] >= 8 * 77 * 16384
    assert lib.scade_mlp_packed_bytes(ctypes.byref(_lib.NetDesc(4, 128, 9, 0, 4))) == 0    # fp32 path only
    st = lib.scade_raw2outputs(None, None, None, 3, None, 4, 8, None, None, None, None, None, None)
    assert st == 1 and b"raw2outputs" in lib.scade_last_error_string()
    st = lib.scade_sample_pdf(None, None, 4, 1, 8, None, 0, None, None, None)
    assert st == 1


def test_cpu_tensors_are_rejected():
    import torch
    from scade_b200 import functional as F_
    with pytest.raises(_lib.ScadeError):
        F_.raw2outputs(torch.zeros(2, 4, 4), torch.zeros(2, 4), torch.ones(2, 3))


def test_stash_image_helpers_roundtrip():
    """The numpy model of the training-stash image (tests/util.py) is self-consistent: swizzle <-> unswizzle and the
    sign-mask bit order (used by the GPU tests to decode what the kernels stashed)."""
    from tests.util import decode_sign_mask, sign_mask_words, stash_swizzle, stash_unswizzle
    rng = np.random.default_rng(0)
    x = rng.standard_normal((256, 128)).astype(np.float16).astype(np.float32)
    img = stash_swizzle(x)
    assert img.size == 2 * 2 * 16384
    np.testing.assert_array_equal(stash_unswizzle(img, 0, 2, 2), x)
    # element (row 9, col 3) of chunk 0: piece 0 ^ (9 & 7) = 1 of row 9
    assert img[9 * 128 + (1 << 4) + 3 * 2:][:2].view(np.float16)[0] == np.float16(x[9, 3])
    neg = rng.random((5, 128)) < 0.5
    np.testing.assert_array_equal(decode_sign_mask(sign_mask_words(neg)), neg)
    one = np.zeros((1, 32), bool)
    one[0, 5] = True                                 # element 5 = 4*1 + 1 -> s = 1, g = 1 -> bit 31 - 8 - 1
    assert sign_mask_words(one)[0, 0] == np.uint32(1 << 22)


def test_cpp_host_example_builds_against_the_header():
    """examples/render_cabi.cpp is a host with no Python and no torch: it must compile against include/scade_b200.h and link
    against the in-tree library (the GPU test runs it)."""
    import os
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    _lib.load()
    r = subprocess.run(["sh", os.path.join(root, "examples", "build.sh")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert os.path.exists(os.path.join(root, "examples", "render_cabi"))


def test_packed_stream_size_follows_the_stage_plan():
    """scade_mlp_packed_bytes = (forward stages + dgrad stages) x 16 KB + the fp32 tail, for every depth / skip the tensor-core
    kernel accepts; shapes it does not handle report 0 (the Python side then raises instead of falling back)."""
    lib = _lib.load()
    tail, stage = 17408, 16384
    for D in range(2, 9):
        for skip in (-1, 0, 2, 4):
            if skip >= D - 1 and skip != -1:
                continue                                   # a skip into the last layer is outside the kernel's plan
            fwd = 2 + sum(10 if (i - 1 == skip) else 8 for i in range(1, D)) + 8 + 6
            bwd = 4 + 8 + 8 * (D - 1)
            got = lib.scade_mlp_packed_bytes(ctypes.byref(_lib.NetDesc(D, 256, 9, 0, skip)))
            assert got == (fwd + bwd) * stage + tail, (D, skip, got)
    assert lib.scade_mlp_packed_bytes(ctypes.byref(_lib.NetDesc(8, 128, 9, 0, 4))) == 0       # width 128: fp32 path only
    assert lib.scade_mlp_packed_bytes(ctypes.byref(_lib.NetDesc(8, 256, 10, 0, 4))) == 0      # 63 encoding columns do not fit
    assert lib.scade_mlp_packed_bytes(ctypes.byref(_lib.NetDesc(8, 256, 9, 2, 4))) == 0       # encoded view directions


def test_control_warps_stay_inside_their_registers():
    """The TMA producer, the MMA issuer and the compositor warp run after `setmaxnreg.dec 32` (csrc/mlp_tc.cu regs_control):
    no instruction of that region or of the local subroutines it calls may name a register above R31 (tools/sass_regions.py)."""
    import shutil
    import subprocess
    import sys
    _lib.load()
    obj = os.path.join(ROOT, "scade_b200", "_lib", "mlp_tc.o")
    if shutil.which("cuobjdump") is None or not os.path.exists(obj):
        pytest.skip("needs cuobjdump and the object file of an in-tree build")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_regions.py"), obj], capture_output=True, text=True, check=True).stdout
    regions = [int(m) for m in re.findall(r"control region \[[^)]*\): max R(\d+)", out)]
    assert len(regions) >= 3 and max(regions) <= 31, out
    comp = out.split("nerf_mlp_tc_pp_kernelILb0ELb1")[1].split("_ZN5scade")[0]
    calls = [int(m) for m in re.findall(r"call target [^:]*: \d+ instructions, max R(\d+)", comp)]
    assert calls and max(calls) <= 31, out


def _composite_plan(lib, S, N, n_sms):
    c, it = ctypes.c_int(), ctypes.c_int()
    assert lib.scade_mlp_composite_plan(S, N, n_sms, ctypes.byref(c), ctypes.byref(it)) == 0
    return c.value, it.value


@pytest.mark.parametrize("n_sms", [148, 132, 8])
def test_fused_compositing_partition(n_sms):
    """How the fused network + compositing kernel splits the points (scade_mlp_composite_plan, the arithmetic its launch uses):
    never more SM pairs than the device has; every point is covered; in chain mode (rays that straddle 128-point tiles, e.g.
    64 + 128 = 192 samples) every CTA's contiguous range starts and ends on a ray boundary, so no ray is split between two
    compositor warps; and the padding stays below one period of steps per CTA."""
    import math
    lib = _lib.load()
    rng = np.random.default_rng(5)
    for S in (32, 64, 96, 128, 160, 192, 224, 256, 320, 384, 512, 1024, 8192):
        for N in [1, 2, 3, 5, 37, 300, 1024, 4096, 32768, 307200] + [int(v) for v in rng.integers(1, 50000, 6)]:
            clusters, iters = _composite_plan(lib, S, N, n_sms)
            P = N * S
            cta_steps = -(-P // 256)
            assert 1 <= clusters <= n_sms // 2
            if S in (32, 64, 128, 256):
                assert iters == 0 and clusters == min((cta_steps + 1) // 2, n_sms // 2)
                continue
            period = S // math.gcd(S, 256)
            assert iters > 0 and iters % period == 0                      # (iters * 256) % S == 0: ranges end on ray boundaries
            assert (iters * 256) % S == 0
            assert 2 * clusters * iters * 256 >= P                        # covered
            assert 2 * (clusters - 1) * iters * 256 < P                   # no idle cluster
            assert iters < -(-cta_steps // (2 * (n_sms // 2))) + period   # at most one period of padding per CTA
    c, it = ctypes.c_int(), ctypes.c_int()
    assert lib.scade_mlp_composite_plan(200, 10, n_sms, ctypes.byref(c), ctypes.byref(it)) != 0      # not a multiple of 32


@pytest.mark.parametrize("retraw", [False, True])
def test_composite_buffers_layout(retraw):
    """functional.composite_buffers: one allocation for the outputs of scade_mlp_forward_rays_composite (raw first, 16-byte
    aligned as the C entry point requires), the reference's shapes (RS:556-560), and a workspace of the size the library asks for."""
    import torch
    from scade_b200 import functional as F_, nerf_helpers as NH
    net = NH.NeRF(D=8, W=256, input_ch=57, input_ch_views=3, output_ch=5, skips=[4], use_viewdirs=True, precision="tc_f16")
    h = net.handle()
    N, S = 37, 192
    b = F_.composite_buffers(h, N, S, "tc_f16", retraw=retraw, device=torch.device("cpu"))
    assert b["weights"].shape == (N, S) and b["rgb"].shape == (N, 3)
    assert b["disp"].shape == b["acc"].shape == b["depth"].shape == (N,)
    assert (b["raw"] is None) == (not retraw)
    if retraw:
        assert b["raw"].shape == (N, S, 4) and b["raw"].data_ptr() % 16 == 0
    ptrs = sorted((t.data_ptr(), t.numel() * 4) for k, t in b.items() if k != "ws" and t is not None)
    for (p0, n0), (p1, _) in zip(ptrs, ptrs[1:]):
        assert p0 + n0 <= p1                                              # views of one buffer, no overlap
    assert b["ws"].numel() == h.workspace_bytes(N * S, _lib.PREC_TC_F16, 0) >= 148 * 5120 and b["ws"].data_ptr() % 16 == 0

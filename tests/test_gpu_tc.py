"""Pins the tcgen05 descriptor conventions the bf16 MLP kernels rely on (ddnerf_b200/csrc/tc.cuh):
operand tile images are built here byte by byte, multiplied by one CTA through the C ABI
(`ddnerf_tc_gemm_selftest`) and compared with a float64 matmul of the same bf16 values.
Tolerance: fp32 accumulation of bf16 products, 2e-5 relative to the row/column norms."""
import ctypes

import numpy as np
import pytest
import torch

from ddnerf_b200 import tcimg

pytestmark = pytest.mark.gpu


def _run(a_img, b_img, N, nk16, a_desc, b_desc, idesc, stepping):
    from ddnerf_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    a = torch.from_numpy(a_img).to(dev)
    b = torch.from_numpy(b_img).to(dev)
    d = torch.full((128, N), float("nan"), device=dev)
    st = (ctypes.c_uint32 * 6)(*stepping)
    _lib.check(lib.ddnerf_tc_gemm_selftest(a.data_ptr(), a.numel(), b.data_ptr(), b.numel(), d.data_ptr(), N, nk16,
                                           ctypes.c_uint64(a_desc), ctypes.c_uint64(b_desc), ctypes.c_uint32(idesc), st,
                                           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "tc_gemm_selftest")
    torch.cuda.synchronize()
    return d.cpu().numpy()


def _operands(N, K, seed):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(128, K, generator=g).bfloat16()
    B = torch.randn(N, K, generator=g).bfloat16()
    ref = A.double().numpy() @ B.double().numpy().T
    return A, B, ref


def _check(d, ref):
    scale = np.abs(ref).max()
    err = np.abs(d - ref).max()
    assert np.isfinite(d).all(), "accumulator holds non-finite values"
    assert err <= 2e-5 * scale * 16, f"max err {err:.3e} (scale {scale:.3e})"


@pytest.mark.parametrize("N,K", [(256, 128), (128, 64), (16, 128), (144, 64)])
def test_kmajor_sw128_both(N, K):
    """A [128,K] and B [N,K] K-major, SWIZZLE_128B blocks of 64 columns (activations in place)."""
    A, B, ref = _operands(N, K, 1)
    a_img = tcimg.kmajor_sw128(A)
    b_img = tcimg.kmajor_sw128(B)
    a_desc = tcimg.smem_desc(0, 0, 1024, tcimg.LAYOUT_SW128)
    b_desc = tcimg.smem_desc(0, 0, 1024, tcimg.LAYOUT_SW128)
    d = _run(a_img, b_img, N, K // 16, a_desc, b_desc, tcimg.idesc_bf16(128, N, 0, 0),
             (32, 4, 128 * 128, 32, 4, N * 128))
    _check(d, ref)


@pytest.mark.parametrize("N,K", [(256, 128), (128, 96), (144, 32)])
def test_kmajor_sw128_a_sw64_b(N, K):
    """A SWIZZLE_128B (activation buffer), B = weight stages [N x 32] in SWIZZLE_64B (the ring)."""
    A, B, ref = _operands(N, K, 2)
    Kp = (K + 63) // 64 * 64
    Ap = torch.zeros(128, Kp, dtype=torch.bfloat16)
    Ap[:, :K] = A
    a_img = tcimg.kmajor_sw128(Ap)
    b_img = tcimg.kmajor_sw64(B)
    a_desc = tcimg.smem_desc(0, 0, 1024, tcimg.LAYOUT_SW128)
    b_desc = tcimg.smem_desc(0, 0, 512, tcimg.LAYOUT_SW64)
    d = _run(a_img, b_img, N, K // 16, a_desc, b_desc, tcimg.idesc_bf16(128, N, 0, 0),
             (32, 4, 128 * 128, 32, 2, N * 64))
    _check(d, ref)


@pytest.mark.parametrize("N,K", [(256, 64), (128, 96)])
def test_kmajor_sw64_both(N, K):
    """A = encoded-feature blocks [128 x 32] SWIZZLE_64B (xyz / dir tiles), B SWIZZLE_64B stages."""
    A, B, ref = _operands(N, K, 3)
    a_img = tcimg.kmajor_sw64(A)
    b_img = tcimg.kmajor_sw64(B)
    a_desc = tcimg.smem_desc(0, 0, 512, tcimg.LAYOUT_SW64)
    b_desc = tcimg.smem_desc(0, 0, 512, tcimg.LAYOUT_SW64)
    d = _run(a_img, b_img, N, K // 16, a_desc, b_desc, tcimg.idesc_bf16(128, N, 0, 0),
             (32, 2, 128 * 64, 32, 2, N * 64))
    _check(d, ref)


@pytest.mark.parametrize("N,K", [(256, 64), (128, 128), (256, 128)])
def test_mnmajor_sw128_both(N, K):
    """dW = dZ^T.X: both operands MN-major views of the row-major [rows, features] tile images
    (K = sample rows).  The images are the same bytes the K-major chain reads."""
    g = torch.Generator().manual_seed(4)
    At = torch.randn(K, 128, generator=g).bfloat16()      # [rows, M features]
    Bt = torch.randn(K, N, generator=g).bfloat16()        # [rows, N features]
    ref = At.double().numpy().T @ Bt.double().numpy()
    a_img = tcimg.kmajor_sw128(At)                        # blocks of 64 features, each [K rows x 128 B]
    b_img = tcimg.kmajor_sw128(Bt)
    blk = K * 128
    a_desc = tcimg.smem_desc(0, blk, 1024, tcimg.LAYOUT_SW128)
    b_desc = tcimg.smem_desc(0, blk, 1024, tcimg.LAYOUT_SW128)
    d = _run(a_img, b_img, N, K // 16, a_desc, b_desc, tcimg.idesc_bf16(128, N, 1, 1),
             (2048, 1 << 20, 0, 2048, 1 << 20, 0))
    _check(d, ref)


# ---------------------------------------------------------------------------------------------
# fused bf16 MLP forward (csrc/mlp_tc.cu)
# ---------------------------------------------------------------------------------------------
def _bf(x):
    return x.bfloat16().float()


def mlp_forward_bf16_emulated(params, x):
    """The kernel's arithmetic restated in torch: bf16 weights and activations, fp32 accumulation and
    bias, density from the bf16 fc_feat output, heads from the bf16 view-branch activations."""
    import torch.nn.functional as F
    W = {k: _bf(v) if k.endswith("weight") else v for k, v in params.items()}
    xyz, dirs = _bf(x[..., :96]), _bf(x[..., 96:])
    h = _bf(F.relu(F.linear(xyz, W["layers_xyz.0.weight"], W["layers_xyz.0.bias"])))
    for i in range(1, 8):
        inp = torch.cat((xyz, h), -1) if i == 5 else h
        h = _bf(F.relu(F.linear(inp, W[f"layers_xyz.{i}.weight"], W[f"layers_xyz.{i}.bias"])))
    feat = _bf(F.linear(h, W["fc_feat.weight"], W["fc_feat.bias"]))
    alpha = F.linear(feat, W["fc_alpha.weight"], W["fc_alpha.bias"])
    hd = _bf(F.relu(F.linear(torch.cat((feat, dirs), -1), W["layers_dir.0.weight"], W["layers_dir.0.bias"])))
    out = [F.linear(hd, W["fc_rgb.weight"], W["fc_rgb.bias"]), alpha]
    if "fc_mu_sigma.weight" in W:
        out.append(F.linear(hd, W["fc_mu_sigma.weight"], W["fc_mu_sigma.bias"]))
    return torch.cat(out, -1)


def _tc_forward(depth_head, N, S, kind, seed):
    from oracle import ddnerf_oracle as orc
    from ddnerf_b200 import mlp_tc
    from ddnerf_b200.models import base_architectures as BA
    from ddnerf_b200.rays import synth_rays
    params = orc.init_mlp_params(depth_head, seed=seed)
    ro, rd, rad, near, far = synth_rays(kind, N, seed=seed)
    rays = orc.pack_rays(ro, rd, rad, near, far)
    g = torch.Generator().manual_seed(seed)
    t_vals = orc.sample_first_cycle(rays[:, 7:8], rays[:, 8:9], S, t_rand=torch.rand(N, S + 1, generator=g))
    x = orc.encode_rows(rays, t_vals)
    ref32 = orc.mlp_forward(params, x)
    ref16 = mlp_forward_bf16_emulated(params, x)
    net = (BA.DepthMipNeRFModel if depth_head else BA.MipNeRFModel)(
        hidden_size=256, max_ipe_deg=16, num_encoding_fn_dir=4, include_input_xyz=False, include_input_dir=True)
    net.load_state_dict(params)
    net.to("cuda")
    with torch.no_grad():
        out = mlp_tc.forward_only(net, rays.cuda(), t_vals.cuda())
    torch.cuda.synchronize()
    return out.cpu(), ref16, ref32


@pytest.mark.parametrize("depth_head,N,S,kind", [(False, 8, 32, "blender"), (True, 37, 16, "blender"),
                                                 (True, 300, 32, "ff"), (False, 2048, 64, "360")])
def test_mlp_tc_forward(depth_head, N, S, kind):
    """Raw network outputs: tight against the bf16-emulating restatement (same rounding points;
    residual = accumulation order + rare 1-ulp bf16 flips), loose against the fp32 oracle."""
    out, ref16, ref32 = _tc_forward(depth_head, N, S, kind, seed=5)
    assert torch.isfinite(out).all()
    err16 = (out - ref16).abs().max().item()
    err32 = (out - ref32).abs().max().item()
    print(f"bf16 MLP forward: max|out-emulated| {err16:.3e}  max|out-fp32| {err32:.3e}  (|ref| max {ref32.abs().max():.3f})")
    assert err16 < 2e-3
    assert err32 < 5e-3

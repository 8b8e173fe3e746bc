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

"""Host-side (numpy) builders of the shared-memory tile images and descriptor words the tcgen05
kernels use (mirror of ddnerf_b200/csrc/tc.cuh).  Used by the tests to pin the layouts; the
production packers are CUDA kernels in csrc/mlp_tc.cu."""
import numpy as np
import torch

LAYOUT_NONE, LAYOUT_SW128, LAYOUT_SW64, LAYOUT_SW32 = 0, 2, 4, 6


def smem_desc(addr, lbo, sbo, layout):
    return ((addr >> 4) & 0x3FFF) | (((lbo >> 4) & 0x3FFF) << 16) | (((sbo >> 4) & 0x3FFF) << 32) | (1 << 46) | (layout << 61)


def idesc_bf16(M, N, a_mn_major, b_mn_major):
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def _bf16_bits(x):
    return x.contiguous().view(torch.int16).numpy().view(np.uint16)


def kmajor_sw128(x):
    """x [R, K] bf16, K % 64 == 0 -> uint8 image: K/64 blocks, each R rows of 128 B, 16-byte chunk c
    of row r stored at chunk c ^ (r & 7)."""
    R, K = x.shape
    assert K % 64 == 0
    bits = _bf16_bits(x)
    img = np.zeros((K // 64) * R * 64, dtype=np.uint16)
    r = np.arange(R)[:, None]
    k = np.arange(K)[None, :]
    blk, kk = k // 64, k % 64
    off = blk * (R * 64) + r * 64 + (((kk >> 3) ^ (r & 7)) << 3) + (kk & 7)
    img[off] = bits
    return img.view(np.uint8)


def kmajor_sw64(x):
    """x [R, K] bf16, K % 32 == 0 -> K/32 blocks, each R rows of 64 B, chunk c (0..3) of row r stored
    at c ^ ((r >> 1) & 3)."""
    R, K = x.shape
    assert K % 32 == 0
    bits = _bf16_bits(x)
    img = np.zeros((K // 32) * R * 32, dtype=np.uint16)
    r = np.arange(R)[:, None]
    k = np.arange(K)[None, :]
    blk, kk = k // 32, k % 32
    off = blk * (R * 32) + r * 32 + (((kk >> 3) ^ ((r >> 1) & 3)) << 3) + (kk & 7)
    img[off] = bits
    return img.view(np.uint8)

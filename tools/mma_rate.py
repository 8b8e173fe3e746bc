#!/usr/bin/env python
"""Tensor-pipe rate probe: cycles per tcgen05.mma (M=128 single CTA / M=256 CTA pair) on resident operands."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddnerf_b200 import _lib
from ddnerf_b200.ops import _p, _stream
lib = _lib.load()
buf = torch.zeros(148, device="cuda", dtype=torch.int64)
for pair in (0, 1):
    for N in (256, 128):
        for ce in (0, 2, 4, 8, 16, -8, -16):
            n = 2048
            for _ in range(2):
                buf.zero_()
                _lib.check(lib.ddnerf_tc_mma_rate(pair, N, n, ce, _p(buf), _stream()), "rate")
                torch.cuda.synchronize()
            v = buf[buf > 0].double()
            print(f"pair={pair} N={N} commit_every={ce}: {v.mean().item() / n:.1f} cycles/MMA (min {v.min().item() / n:.1f}, max {v.max().item() / n:.1f}), "
                  f"{len(v)} issuing CTAs")

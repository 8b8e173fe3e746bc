#!/usr/bin/env python
"""Is the weight ring of the chain kernels latency- or bandwidth-bound?  Runs the dX chain on fewer and fewer CTAs and
prints the issuer's cycles per work unit: constant = per-SM latency, falling with fewer CTAs = shared L2 bandwidth."""
import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddnerf_b200 import _lib, mlp_tc
from ddnerf_b200.models import base_architectures as BA
from ddnerf_b200.ops import _p, _stream
lib = _lib.load()
pair = int(sys.argv[1]) if len(sys.argv) > 1 else 1
lib.ddnerf_mlp_tc_set_pair_mode(pair)
rows = 4096 * 128
net = BA.MipNeRFModel(hidden_size=256, max_ipe_deg=16, num_encoding_fn_dir=4, include_input_xyz=False, include_input_dir=True).to("cuda")
st = mlp_tc._state(net); st.refresh()
mask = torch.randint(0, 255, (lib.ddnerf_mlp_tc_mask_save_bytes(rows),), device="cuda", dtype=torch.uint8)
dz = torch.empty(lib.ddnerf_mlp_tc_act_save_bytes(rows), device="cuda", dtype=torch.uint8)
gout = torch.randn(rows, 4, device="cuda")
buf = torch.zeros(148 * 8, device="cuda", dtype=torch.int64)
for ctas in (148,):
    for prof in (0, 1):
        lib.ddnerf_mlp_tc_set_profile_buffer(_p(buf) if prof else None)
        buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lib.ddnerf_mlp_tc_backward_dx(_p(st.wimg), _p(st.bias), _p(gout), rows, 4, _p(mask), _p(dz), ctas, _stream())
        e0.record()
        _lib.check(lib.ddnerf_mlp_tc_backward_dx(_p(st.wimg), _p(st.bias), _p(gout), rows, 4, _p(mask), _p(dz), ctas, _stream()), "dx")
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if prof:
            b = buf.view(148, 8)[:ctas].double()
            b = b[0::2] if pair else b
            units = (rows / 512) / (ctas / 2) if pair else (rows / 256) / ctas
            tot, t_act, t_stage = [b[:, i].mean().item() for i in range(3)]
            e_wait, e_busy, n_e, e_pre, e_work = [b[:, i].mean().item() for i in range(3, 8)]
            print(f"pair={pair} ctas={ctas:3d}: {ms:.3f} ms | issuer {tot / units:.0f} cyc per unit ({units:.1f} units), ring wait {100 * t_stage / tot:.1f}%, epilogue wait {100 * t_act / tot:.1f}%"
                  f" | epilogue: n={n_e:.0f} wait {e_wait / n_e:.0f} busy {e_busy / n_e:.0f} (pre {e_pre / n_e:.0f} work {e_work / n_e:.0f})", flush=True)
lib.ddnerf_mlp_tc_set_profile_buffer(None)

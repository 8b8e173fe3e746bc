#!/usr/bin/env python
"""Micro-benchmark of the fused bf16 MLP kernels (CUDA events, after warm-up).

    python tools/bench_mlp_tc.py [--rows 1048576] [--iters 10]
"""
import argparse
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=8192)
    ap.add_argument("--samples", type=int, default=128)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--save", action="store_true", help="training forward (store activations + masks)")
    ap.add_argument("--prof", action="store_true", help="print the chain kernels' per-role cycle counters")
    ap.add_argument("--pair", type=int, default=-1, help="1: CTA-pair chain kernels, 0: single-CTA kernels, -1: library default")
    args = ap.parse_args()
    from oracle import ddnerf_oracle as orc
    from ddnerf_b200 import _lib, mlp_tc
    from ddnerf_b200.models import base_architectures as BA
    from ddnerf_b200.ops import _p, _stream
    from ddnerf_b200.rays import synth_rays
    lib = _lib.load()
    lib.ddnerf_mlp_tc_set_pair_mode(args.pair)
    N, S = args.rays, args.samples
    ro, rd, rad, near, far = synth_rays("blender", N, seed=1)
    rays = orc.pack_rays(ro, rd, rad, near, far).cuda()
    t_vals = orc.sample_first_cycle(rays[:, 7:8].cpu(), rays[:, 8:9].cpu(), S).cuda()
    net = BA.MipNeRFModel(hidden_size=256, max_ipe_deg=16, num_encoding_fn_dir=4, include_input_xyz=False, include_input_dir=True)
    net.to("cuda")
    st = mlp_tc._state(net)
    st.refresh()
    rows = N * S
    out = torch.empty(rows, 4, device="cuda")
    act = mask = None
    if args.save:
        act = torch.empty(lib.ddnerf_mlp_tc_act_save_bytes(rows), device="cuda", dtype=torch.uint8)
        mask = torch.empty(lib.ddnerf_mlp_tc_mask_save_bytes(rows), device="cuda", dtype=torch.uint8)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    res = {}
    for name, fn in (("encode", lambda: mlp_tc.encode_img(rays, t_vals)),):
        for _ in range(3):
            img = fn()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(args.iters):
            img = fn()
        e1.record()
        torch.cuda.synchronize()
        res[name + "_ms"] = e0.elapsed_time(e1) / args.iters

    def fwd():
        _lib.check(lib.ddnerf_mlp_tc_forward(_p(st.wimg), _p(st.bias), _p(img), rows, 4, _p(out), _p(act), _p(mask), _stream()), "fwd")

    for _ in range(3):
        fwd()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(args.iters):
        fwd()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    flops = 2.0 * 610304 * rows
    res.update(pair=args.pair, rows=rows, fwd_ms=ms, fwd_tflops=flops / ms / 1e9, save=bool(args.save))

    # the model path: ONE launch, the chain kernel's own encoder warps (no standalone encode)
    scratch = torch.empty(lib.ddnerf_mlp_tc_enc_scratch_bytes(), device="cuda", dtype=torch.uint8)
    img2 = torch.empty_like(img) if args.save else None

    def fwd_rays():
        _lib.check(lib.ddnerf_mlp_tc_forward_rays(_p(st.wimg), _p(st.bias), _p(rays), _p(t_vals), N, S, 0, 4, _p(out), _p(img2),
                                                  None if args.save else _p(scratch), _p(act), _p(mask), _stream()), "fwd_rays")

    for _ in range(3):
        fwd_rays()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(args.iters):
        fwd_rays()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    res.update(fwd_rays_ms=ms, fwd_rays_tflops=flops / ms / 1e9, encode_plus_fwd_ms=res["encode_ms"] + res["fwd_ms"])
    if args.save:
        dz = torch.empty_like(act)
        gout = torch.randn(rows, 4, device="cuda")
        pairs = net._param_pairs()
        gws = [torch.zeros_like(w) for w, _ in pairs]
        gbs = [torch.zeros_like(b) for _, b in pairs]
        from ddnerf_b200.ops import _ptr_table
        table = _ptr_table(gws, gbs)

        def dx():
            _lib.check(lib.ddnerf_mlp_tc_backward_dx(_p(st.wimg), _p(st.bias), _p(gout), rows, 4, _p(mask), _p(dz), 0, _stream()), "dx")

        def dw():
            _lib.check(lib.ddnerf_mlp_tc_backward_dw(_p(act), _p(dz), _p(img), _p(gout), ctypes.byref(table), rows, 4, 0, _stream()), "dw")

        for name, fn, macs in (("dx", dx, 557696), ("dw", dw, 610304)):
            for _ in range(2):
                fn()
            e0, e1 = ev(), ev()
            e0.record()
            for _ in range(args.iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            res[name + "_ms"] = ms
            res[name + "_tflops"] = 2.0 * macs * rows / ms / 1e9
        res["dw_gbs"] = (2 * act.numel()) / res["dw_ms"] / 1e6
        tot = res["fwd_ms"] + res["dx_ms"] + res["dw_ms"]
        res["train_tflops"] = 2.0 * (610304 * 2 + 557696) * rows / tot / 1e9
    print(json.dumps(res))
    if args.prof:
        buf = torch.zeros(148 * 8, device="cuda", dtype=torch.int64)
        lib.ddnerf_mlp_tc_set_profile_buffer(_p(buf))
        runs = [("fwd", fwd), ("fwd_rays", fwd_rays)] + ([("dx", dx)] if args.save else [])
        for name, fn in runs:
            buf.zero_()
            fn()
            torch.cuda.synchronize()
            if args.pair != 0:
                pb_ = buf.view(148, 8).double()[1::2]
                pt, pe, pn = [pb_[:, i].mean().item() for i in range(3)]
                print(f"{name}: peer producer total {pt:.0f} cyc | waiting for free slots {100 * pe / max(pt, 1):.1f}% | for the encoder {100 * pn / max(pt, 1):.1f}%")
            b = buf.view(148, 8).double()[0::2] if args.pair != 0 else buf.view(148, 8).double()
            tot, t_act, t_stage, e_wait, e_busy, n, e_pre, e_work = [b[:, i].mean().item() for i in range(8)]
            if os.environ.get("DDNERF_TC_PROF_ISSUER"):
                print(f"{name}: issuer total {tot:.0f} | act waits {100*t_act/tot:.1f}% | T0 passes {100*t_stage/tot:.1f}% | waits for own stages {100*e_wait/tot:.1f}% | waits for the peer's {100*e_busy/tot:.1f}%")
                continue
            print(f"{name}: issuer total {tot:.0f} cyc | waits: epilogue {100 * t_act / tot:.1f}% ring {100 * t_stage / tot:.1f}% | "
                  f"epilogue busy {e_busy / max(n, 1):.0f} cyc each (store drain + barrier {e_pre / max(n, 1):.0f}, accumulator -> act buffer "
                  f"{e_work / max(n, 1):.0f}), waiting on MMA {e_wait / max(n, 1):.0f} cyc each, n={n:.0f}")
        lib.ddnerf_mlp_tc_set_profile_buffer(None)
        if args.save:
            import collections
            pb = torch.zeros(4 * 480, device="cuda", dtype=torch.int64)
            lib.ddnerf_mlp_tc_dw_set_profile_buffer(_p(pb))
            dw()
            torch.cuda.synchronize()
            lib.ddnerf_mlp_tc_dw_set_profile_buffer(None)
            rowsp = pb.view(480, 4).cpu().tolist()
            per = collections.defaultdict(list)
            for op, tiles, cyc, fl in rowsp:
                if tiles > 0:
                    per[op].append((tiles, cyc, fl))
            for op in sorted(per):
                v = per[op]
                big = [x for x in v if x[0] > 20]
                cpt = sum(x[1] for x in big) / max(sum(x[0] for x in big), 1)
                print(f"dw op {op:2d}: items {len(v):3d} tiles {sum(x[0] for x in v):6d} cycles/tile {cpt:8.0f} flush {sum(x[2] for x in v) / len(v):8.0f} cyc/item")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""A/B check of the two chain-kernel variants (single CTA vs CTA pair): forward outputs, saved activations, ReLU masks,
the encoded images and the dZ images must be BIT-identical (same MMA shapes along K, same epilogues).

    python tools/check_pair.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(lib, mode, net_C, rays, t_vals, st, gout):
    from ddnerf_b200 import _lib
    from ddnerf_b200.ops import _p, _stream
    lib.ddnerf_mlp_tc_set_pair_mode(mode)
    N, S = t_vals.shape[0], t_vals.shape[1] - 1
    rows = N * S
    C = net_C
    out = torch.zeros(rows, C, device="cuda")
    out_inf = torch.zeros(rows, C, device="cuda")
    act = torch.zeros(lib.ddnerf_mlp_tc_act_save_bytes(rows), device="cuda", dtype=torch.uint8)
    mask = torch.zeros(lib.ddnerf_mlp_tc_mask_save_bytes(rows), device="cuda", dtype=torch.uint8)
    img = torch.zeros(lib.ddnerf_mlp_tc_enc_bytes(rows), device="cuda", dtype=torch.uint8)
    scratch = torch.zeros(lib.ddnerf_mlp_tc_enc_scratch_bytes(), device="cuda", dtype=torch.uint8)
    dz = torch.zeros_like(act)
    _lib.check(lib.ddnerf_mlp_tc_forward_rays(_p(st.wimg), _p(st.bias), _p(rays), _p(t_vals), N, S, 0, C, _p(out), _p(img), None,
                                              _p(act), _p(mask), _stream()), "fwd_rays train")
    _lib.check(lib.ddnerf_mlp_tc_forward_rays(_p(st.wimg), _p(st.bias), _p(rays), _p(t_vals), N, S, 0, C, _p(out_inf), None,
                                              _p(scratch), None, None, _stream()), "fwd_rays inference")
    out_img = torch.zeros(rows, C, device="cuda")
    _lib.check(lib.ddnerf_mlp_tc_forward(_p(st.wimg), _p(st.bias), _p(img), rows, C, _p(out_img), None, None, _stream()), "fwd image")
    _lib.check(lib.ddnerf_mlp_tc_backward_dx(_p(st.wimg), _p(st.bias), _p(gout), rows, C, _p(mask), _p(dz), 0, _stream()), "dx")
    torch.cuda.synchronize()
    return dict(out=out, out_inf=out_inf, out_img=out_img, act=act, mask=mask, img=img, dz=dz)


def main():
    import hashlib
    import json
    digests = {}
    from oracle import ddnerf_oracle as orc
    from ddnerf_b200 import _lib, mlp_tc
    from ddnerf_b200.models import base_architectures as BA
    from ddnerf_b200.rays import synth_rays
    lib = _lib.load()
    bad = 0
    for (N, S, dd) in ((3, 128, False), (5, 256, True), (7, 128, False), (1024, 32, True), (4096, 128, False)):
        torch.manual_seed(N)
        ro, rd, rad, near, far = synth_rays("blender", N, seed=1)
        rays = orc.pack_rays(ro, rd, rad, near, far).cuda()
        t_vals = orc.sample_first_cycle(rays[:, 7:8].cpu(), rays[:, 8:9].cpu(), S).cuda()
        cls = BA.DepthMipNeRFModel if dd else BA.MipNeRFModel
        net = cls(hidden_size=256, max_ipe_deg=16, num_encoding_fn_dir=4, include_input_xyz=False, include_input_dir=True)
        net.to("cuda")
        C = 6 if dd else 4
        st = mlp_tc._state(net)
        st.refresh()
        gout = torch.randn(N * S, C, device="cuda")
        a = run(lib, 0, C, rays, t_vals, st, gout)
        b = run(lib, 1, C, rays, t_vals, st, gout)
        for k in a:
            same = torch.equal(a[k], b[k])
            if not same:
                bad += 1
                x, y = a[k], b[k]
                if x.dtype == torch.uint8:
                    nd = (x != y).sum().item()
                    first = (x != y).nonzero()[0].item()
                    print(f"  N={N} S={S} C={C} {k}: {nd} of {x.numel()} bytes differ, first at {first}")
                else:
                    print(f"  N={N} S={S} C={C} {k}: max abs diff {(x - y).abs().max().item():.3e}, nan {torch.isnan(y).sum().item()}")
        for k in b:        # digests of the written parts (uninitialised bytes of the image buffers are zero: the buffers start zeroed)
            digests[f"{N}x{S}x{C}:{k}"] = hashlib.sha1(b[k].cpu().numpy().tobytes()).hexdigest()
        print(f"N={N} S={S} C={C}: " + ("identical" if all(torch.equal(a[k], b[k]) for k in a) else "DIFFERENT"), flush=True)
    if len(sys.argv) > 1:          # digests for comparing two processes (e.g. DDNERF_TC_REG_SAVES=0 against the default)
        json.dump(digests, open(sys.argv[1], "w"), indent=0)
    print("check_pair:", "OK" if bad == 0 else f"{bad} mismatches")
    sys.exit(0 if bad == 0 else 1)


if __name__ == "__main__":
    main()

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


@pytest.mark.parametrize("N,K", [(64, 64), (32, 128), (96, 64)])
def test_mnmajor_sw128_a_sw64_b(N, K):
    """dW of the encoded inputs: A = dZ tile (MN-major SWIZZLE_128B), B = [rows x 32] SWIZZLE_64B
    encoder blocks read MN-major (LBO = block pitch, SBO = 8 rows of 64 B, 1 KB per k16 step)."""
    g = torch.Generator().manual_seed(6)
    At = torch.randn(K, 128, generator=g).bfloat16()
    Bt = torch.randn(K, N, generator=g).bfloat16()
    ref = At.double().numpy().T @ Bt.double().numpy()
    a_img = tcimg.kmajor_sw128(At)
    b_img = tcimg.kmajor_sw64(Bt)
    a_desc = tcimg.smem_desc(0, K * 128, 1024, tcimg.LAYOUT_SW128)
    b_desc = tcimg.smem_desc(0, K * 64, 512, tcimg.LAYOUT_SW64)
    d = _run(a_img, b_img, N, K // 16, a_desc, b_desc, tcimg.idesc_bf16(128, N, 1, 1),
             (2048, 1 << 20, 0, 1024, 1 << 20, 0))
    _check(d, ref)


# ---------------------------------------------------------------------------------------------
# fused bf16 MLP backward (csrc/mlp_tc.cu dX chain + csrc/mlp_tc_dw.cu)
# ---------------------------------------------------------------------------------------------
class _GradRound(torch.autograd.Function):
    """identity whose cotangent is rounded to bf16 (where the dX chain stores dZ)"""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


class _BfSTE(torch.autograd.Function):
    """bf16 rounding with a straight-through gradient"""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


def mlp_train_bf16_emulated(params, x):
    """Forward AND backward arithmetic of the tensor-core kernels restated with torch autograd: bf16
    weights / activations / dZ, fp32 accumulation; the ReLU masks come from the same rounded forward,
    which is what makes a tight comparison possible (against fp32 autograd every unit whose
    pre-activation changes sign under bf16 rounding contributes an O(1) relative difference)."""
    import torch.nn.functional as F
    W = {k: (v + (_bf(v) - v).detach()) if k.endswith("weight") else v for k, v in params.items()}
    xyz, dirs = _bf(x[..., :96]), _bf(x[..., 96:])

    def layer(inp, name, relu=True):
        z = _GradRound.apply(F.linear(inp, W[name + ".weight"], W[name + ".bias"]))
        return _BfSTE.apply(F.relu(z) if relu else z)

    h = layer(xyz, "layers_xyz.0")
    for i in range(1, 8):
        h = layer(torch.cat((xyz, h), -1) if i == 5 else h, f"layers_xyz.{i}")
    feat = layer(h, "fc_feat", relu=False)
    alpha = F.linear(feat, W["fc_alpha.weight"], W["fc_alpha.bias"])
    hd = layer(torch.cat((feat, dirs), -1), "layers_dir.0")
    out = [F.linear(hd, W["fc_rgb.weight"], W["fc_rgb.bias"]), alpha]
    if "fc_mu_sigma.weight" in W:
        out.append(F.linear(hd, W["fc_mu_sigma.weight"], W["fc_mu_sigma.bias"]))
    return _GradRound.apply(torch.cat(out, -1))


@pytest.mark.parametrize("depth_head,N,S,kind", [(False, 8, 32, "blender"), (True, 37, 16, "blender"),
                                                 (True, 600, 32, "360"), (False, 1024, 64, "ff")])
def test_mlp_tc_backward(depth_head, N, S, kind):
    """Parameter gradients of sum(out * cotangent) against the bf16-emulating autograd restatement
    (same rounding points and ReLU masks): per tensor, cosine similarity above 0.9995 and max element error
    below 10 % of the tensor's largest gradient (isolated elements move when one of the few ReLU units
    whose sign differs between the two accumulation orders flips; a layout or indexing bug gives O(1)
    everywhere).  The distance to
    fp32 autograd of the oracle is printed: it is dominated by ReLU units flipping under bf16 rounding."""
    from oracle import ddnerf_oracle as orc
    from ddnerf_b200 import mlp_tc
    from ddnerf_b200.models import base_architectures as BA
    from ddnerf_b200.rays import synth_rays
    seed = 11
    params = orc.init_mlp_params(depth_head, seed=seed)
    ro, rd, rad, near, far = synth_rays(kind, N, seed=seed)
    rays = orc.pack_rays(ro, rd, rad, near, far)
    g = torch.Generator().manual_seed(seed)
    t_vals = orc.sample_first_cycle(rays[:, 7:8], rays[:, 8:9], S, t_rand=torch.rand(N, S + 1, generator=g))
    C = 6 if depth_head else 4
    ct = torch.randn(N * S, C, generator=g)
    x = orc.encode_rows(rays, t_vals)
    p32 = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    (orc.mlp_forward(p32, x) * ct).sum().backward()
    p16 = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    (mlp_train_bf16_emulated(p16, x) * ct).sum().backward()

    net = (BA.DepthMipNeRFModel if depth_head else BA.MipNeRFModel)(
        hidden_size=256, max_ipe_deg=16, num_encoding_fn_dir=4, include_input_xyz=False, include_input_dir=True)
    net.load_state_dict(params)
    net.to("cuda")
    out = mlp_tc.mlp_bf16(net, rays.cuda(), t_vals.cuda())
    assert out.requires_grad
    (out * ct.cuda()).sum().backward()
    torch.cuda.synchronize()
    worst, bad = 0.0, []
    for k, p in net.named_parameters():
        got = p.grad.cpu()
        assert torch.isfinite(got).all(), k
        ref = p16[k].grad
        rel = ((got - ref).abs().max() / ref.abs().max().clamp(min=1e-12)).item()
        cos = torch.nn.functional.cosine_similarity(got.flatten().double(), ref.flatten().double(), dim=0).item()
        r32 = p32[k].grad
        rel32 = ((got - r32).abs().max() / r32.abs().max().clamp(min=1e-12)).item()
        worst = max(worst, rel)
        print(f"  {k:24s} vs bf16-emulated: rel {rel:.3e} cos {cos:.6f} | vs fp32: rel {rel32:.3e}")
        if not (rel < 0.1 and cos > 0.9995):
            bad.append((k, rel, cos))
    print(f"bf16 MLP backward: worst per-tensor max error {worst:.3e} of the tensor's max |grad|")
    assert not bad, bad


@pytest.mark.parametrize("pname", ["config_blender", "config_blender_mipnerf", "config_360"])
def test_train_step_bf16_vs_oracle(pname):
    """One train step through the model API with the bf16 tensor-core MLP (forward + backward)
    against the fp32 oracle on the same weights / rays / random draws: rendered rgb and depth within
    the 5e-3 budget BASELINE.json gives the bf16 mode, loss within 5e-3, gradient direction within
    cos > 0.97 of fp32 autograd."""
    from oracle import ddnerf_oracle as orc
    from ddnerf_b200.config import preset
    from ddnerf_b200.models import models as M
    from ddnerf_b200.rays import synth_rays
    cfg, kind = preset(pname, num_coarse=32, num_fine=32)
    is_dd = cfg.nerf.type == "DDNerfModel"
    N, s0, s1 = 512, 32, 32
    ro, rd, rad, near, far = synth_rays(kind, N, seed=3)
    g = torch.Generator().manual_seed(0)
    target = torch.rand(N, 3, generator=g)
    rnd = dict(t_rand=torch.rand(N, s0 + 1, generator=g), noise0=torch.randn(N, s0, generator=g),
               u_rand=torch.rand(N, s1 + 1, generator=g), noise1=torch.randn(N, s1, generator=g))
    pc = orc.init_mlp_params(is_dd, seed=7)
    pf = orc.init_mlp_params(False, seed=8) if is_dd else None
    tp = cfg.train_params
    ocfg = orc.PathConfig(model=cfg.nerf.type, near=near, far=far, num_coarse=s0, num_fine=s1, perturb=True,
                          noise_std=cfg.nerf.train.radiance_field_noise_std, blender=cfg.dataset.type.lower() == "blender",
                          pdf_padding=tp.pdf_padding, gaussian_smooth_factor=tp.gaussian_smooth_factor,
                          dist_reg_coeficient=tp.dist_reg_coeficient, loss_coeficients=tp.loss_coeficients,
                          dp_coeficient=tp.dp_coeficient)
    rays = orc.pack_rays(ro, rd, rad, near, far)
    loss_ref, out_ref, gc_ref, gf_ref = orc.train_step(ocfg, pc, pf, rays, target, rnd)

    dev = torch.device("cuda:0")
    model = getattr(M, cfg.nerf.type)(cfg)
    model.coarse.load_state_dict(pc)
    model.coarse.mlp_mode = "bf16"
    if is_dd:
        model.fine.load_state_dict(pf)
        model.fine.mlp_mode = "bf16"
    model.to(dev)
    model.randoms = {k: v.to(dev) for k, v in rnd.items()}
    out = model.run_iter(ro.to(dev), rd.to(dev), rad.to(dev), mode="train", rgb_target=target.to(dev))
    tgt = target.to(dev)
    loss = sum(tp.loss_coeficients[j] * torch.nn.functional.mse_loss(out[j]["rgb"], tgt) for j in range(2))
    if is_dd:
        loss = loss + tp.dp_coeficient * out[1]["dp_loss"].mean()
    loss.backward()
    torch.cuda.synchronize()
    for j in range(2):
        e_rgb = (out[j]["rgb"].detach().cpu() - out_ref[j]["rgb"]).abs().max().item()
        e_dep = (out[j]["depth"].detach().cpu() - out_ref[j]["depth"]).abs().max().item()
        print(f"{pname} pass {j}: |rgb| {e_rgb:.2e} |depth| {e_dep:.2e}")
        assert e_rgb < 5e-3 and e_dep < 5e-3
    assert abs(loss.item() - loss_ref.item()) < 5e-3
    for ref_g, net in ((gc_ref, model.coarse), (gf_ref, model.fine if is_dd else None)):
        if net is None:
            continue
        for k, p in net.named_parameters():
            ref, got = ref_g[k], p.grad.cpu()
            rel = ((got - ref).abs().max() / ref.abs().max().clamp(min=1e-12)).item()
            cos = torch.nn.functional.cosine_similarity(got.flatten().double(), ref.flatten().double(), dim=0).item()
            print(f"  {k:24s} rel {rel:.3e} cos {cos:.5f}")
            # fp32 reference: ReLU units that flip under bf16 rounding bound this from below (see
            # test_mlp_tc_backward); direction must agree
            assert cos > 0.97 and rel < 0.35, (k, rel, cos)


@pytest.mark.parametrize("pname", ["config_blender_mipnerf", "config_360"])
def test_trainer_cuda_graph_matches_eager(pname):
    """Trainer(use_graph=True) replays one captured CUDA graph per iteration; with the random draws injected it
    must walk exactly the trajectory of the eager trainer (same kernels, same order): weights after 6
    iterations agree to fp32 round-off of the atomically accumulated gradients."""
    from oracle import ddnerf_oracle as orc
    from ddnerf_b200.config import preset
    from ddnerf_b200.models import models as M
    from ddnerf_b200.rays import synth_rays
    from ddnerf_b200.trainer import Trainer
    dev = torch.device("cuda:0")
    N, s0, s1 = 512, 32, 32
    ro, rd, rad, near, far = synth_rays("blender" if "blender" in pname else "360", N, seed=4)
    g = torch.Generator().manual_seed(1)
    target = torch.rand(N, 3, generator=g)
    rnd = dict(t_rand=torch.rand(N, s0 + 1, generator=g), noise0=torch.randn(N, s0, generator=g),
               u_rand=torch.rand(N, s1 + 1, generator=g), noise1=torch.randn(N, s1, generator=g))
    finals = []
    for use_graph in (False, True):
        cfg, _ = preset(pname, num_coarse=s0, num_fine=s1)
        is_dd = cfg.nerf.type == "DDNerfModel"
        model = getattr(M, cfg.nerf.type)(cfg)
        model.coarse.load_state_dict(orc.init_mlp_params(is_dd, seed=21))
        model.coarse.mlp_mode = "bf16"
        if is_dd:
            model.fine.load_state_dict(orc.init_mlp_params(False, seed=22))
            model.fine.mlp_mode = "bf16"
        model.to(dev)
        model.randoms = {k: v.to(dev) for k, v in rnd.items()}
        tr = Trainer(model, use_graph=use_graph)
        args = [t.to(dev) for t in (ro, rd, rad, target)]
        losses = []
        for _ in range(6):
            loss, mse = tr.step(*args)
            losses.append(loss.item())
        assert use_graph == (tr._graph is not None)
        finals.append((losses, torch.cat([b.flat.clone() for b in tr.buckets]).cpu()))
    (l0, w0), (l1, w1) = finals
    assert max(abs(a - b) for a, b in zip(l0, l1)) < 1e-5, (l0, l1)
    # Adam normalises each gradient element, so an element whose gradient is pure round-off noise can move by a full
    # learning-rate step either way: compare against the step size (lr ~ 5e-6 at these iterations), not ulp
    assert (w0 - w1).abs().max().item() < 2e-4
    assert torch.nn.functional.cosine_similarity((w0 - w0.mean()).double(), (w1 - w1.mean()).double(), dim=0).item() > 0.999999


@pytest.mark.parametrize("pname", ["config_blender_mipnerf", "config_360"])
def test_gradient_sink_equals_autograd_gradients(pname):
    """FlatBucket.install_sink(): the dW kernel accumulating straight into the trainer's flat gradient bucket gives the
    gradients autograd would have left in p.grad (both passes of the shared mip-NeRF network, both DDNeRF networks)."""
    from oracle import ddnerf_oracle as orc
    from ddnerf_b200.config import preset
    from ddnerf_b200.models import models as M
    from ddnerf_b200.rays import synth_rays
    from ddnerf_b200.trainer import FlatBucket
    dev = torch.device("cuda:0")
    N, s0, s1 = 384, 32, 32
    ro, rd, rad, near, far = synth_rays("blender" if "blender" in pname else "360", N, seed=9)
    g = torch.Generator().manual_seed(2)
    target = torch.rand(N, 3, generator=g).to(dev)
    rnd = dict(t_rand=torch.rand(N, s0 + 1, generator=g), noise0=torch.randn(N, s0, generator=g),
               u_rand=torch.rand(N, s1 + 1, generator=g), noise1=torch.randn(N, s1, generator=g))
    grads = []
    for sink in (False, True):
        cfg, _ = preset(pname, num_coarse=s0, num_fine=s1)
        is_dd = cfg.nerf.type == "DDNerfModel"
        model = getattr(M, cfg.nerf.type)(cfg)
        model.coarse.load_state_dict(orc.init_mlp_params(is_dd, seed=21))
        nets = [model.coarse]
        if is_dd:
            model.fine.load_state_dict(orc.init_mlp_params(False, seed=22))
            nets.append(model.fine)
        for net in nets:
            net.mlp_mode = "bf16"
        model.to(dev)
        model.record_distributions = False
        model.randoms = {k: v.to(dev) for k, v in rnd.items()}
        buckets = [FlatBucket(net) for net in nets]
        for b in buckets:
            if sink:
                assert b.install_sink()
            b.begin_step()
        model.train()
        out = model.run_iter(ro.to(dev), rd.to(dev), rad.to(dev), mode="train", rgb_target=target)
        loss = sum(torch.nn.functional.mse_loss(out[j]["rgb"], target) for j in range(2))
        if is_dd:
            loss = loss + 0.1 * out[1]["dp_loss"].mean()
        loss.backward()
        if sink:
            assert all(p.grad is None for b in buckets for p in b.params)
        for b in buckets:
            b.gather_grads()
        grads.append(torch.cat([b.grad.clone() for b in buckets]).cpu())
    a, s = grads
    assert torch.isfinite(s).all() and s.abs().max() > 0
    # same kernels, same operands; only the order of the fp32 atomic accumulation differs
    assert ((a - s).abs().max() / a.abs().max()).item() < 1e-4
    assert torch.nn.functional.cosine_similarity(a.double(), s.double(), dim=0).item() > 0.9999999


def _decode_enc_image(img, rows):
    """Rows of the chain kernels' encoded operand image back to [rows, 96] IPE and [rows, 32] direction features
    (layout: DESIGN.md section 3 -- per 256-row item [T0 xyz0,xyz1 | T1 xyz0,xyz1 | T0 xyz2 | T1 xyz2 | T0 dir | T1 dir],
    blocks of [128 rows x 32 bf16], K-major SWIZZLE_64B)."""
    raw = img.cpu().numpy().view(np.uint16)
    n_items = (rows + 255) // 256
    ipe = np.zeros((n_items * 256, 96), dtype=np.uint16)
    dirs = np.zeros((n_items * 256, 32), dtype=np.uint16)
    r = np.arange(128)
    for it in range(n_items):
        base = it * 32768                                    # uint16 units (64 KB per item)
        for T in range(2):
            for b in range(4):
                off = {0: T * 8192, 1: T * 8192 + 4096, 2: 16384 + T * 4096, 3: 24576 + T * 4096}[b]
                for j in range(4):
                    col = (j ^ ((r >> 1) & 3)) * 8
                    src = base + off + r * 32
                    vals = np.stack([raw[src + col + k] for k in range(8)], 1)        # [128, 8]
                    dst_rows = it * 256 + T * 128 + r
                    if b < 3:
                        ipe[dst_rows, b * 32 + j * 8: b * 32 + j * 8 + 8] = vals
                    else:
                        dirs[dst_rows, j * 8: j * 8 + 8] = vals
    to_f = lambda u: torch.from_numpy((u.astype(np.uint32) << 16).view(np.float32).copy())
    return to_f(ipe)[:rows], to_f(dirs)[:rows]


def _chain_encoder_image(rays, t_vals):
    """The operand image written by the ENCODER WARPS of the forward chain kernel (training forward: the image is kept for
    the weight-gradient kernel), through the C ABI."""
    import ctypes
    from ddnerf_b200 import _lib, mlp_tc
    from ddnerf_b200.models import base_architectures as BA
    from oracle import ddnerf_oracle as orc
    lib = _lib.load()
    net = BA.MipNeRFModel(hidden_size=256, max_ipe_deg=16, num_encoding_fn_dir=4, include_input_xyz=False, include_input_dir=True)
    net.load_state_dict(orc.init_mlp_params(False, seed=1))
    net.to("cuda")
    st = mlp_tc._state(net)
    st.refresh()
    N, S = rays.shape[0], t_vals.shape[1] - 1
    rows = N * S
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    img = torch.full((lib.ddnerf_mlp_tc_enc_bytes(rows),), 0xAB, device="cuda", dtype=torch.uint8)
    out = torch.empty(rows, 4, device="cuda")
    act = torch.empty(lib.ddnerf_mlp_tc_act_save_bytes(rows), device="cuda", dtype=torch.uint8)
    mask = torch.empty(lib.ddnerf_mlp_tc_mask_save_bytes(rows), device="cuda", dtype=torch.uint8)
    _lib.check(lib.ddnerf_mlp_tc_forward_rays(P(st.wimg), P(st.bias), P(rays), P(t_vals), N, S, 0, 4, P(out), P(img), None, P(act),
                                              P(mask), None), "mlp_tc_forward_rays")
    torch.cuda.synchronize()
    return img


@pytest.mark.parametrize("source", ["chain_kernel_encoder_warps", "standalone_kernel"])
@pytest.mark.parametrize("kind,N,S", [("blender", 37, 16), ("ff", 64, 32), ("360", 9, 128), ("blender", 700, 64)])
def test_encode_img_vs_oracle(kind, N, S, source):
    """K2 in bf16 mode: every feature of the operand image within one bf16 step of the oracle's fp32 encoding
    (integrated positional encoding, math_utils.py:112-166, and the view-direction encoding, nerf_helpers.py:127-171) --
    for the image the forward chain kernel's own encoder warps write (the model path; 700 x 64 rows = 175 work items, more
    than one per CTA, so the double-buffer hand-shake between encoder, producer and MMA warps is exercised) and for the
    standalone encoder kernel."""
    from oracle import ddnerf_oracle as orc
    from ddnerf_b200 import mlp_tc
    from ddnerf_b200.rays import synth_rays
    ro, rd, rad, near, far = synth_rays(kind, N, seed=11)
    rays = orc.pack_rays(ro, rd, rad, near, far)
    g = torch.Generator().manual_seed(5)
    t = orc.sample_first_cycle(torch.full((N, 1), near), torch.full((N, 1), far), S, False, torch.rand(N, S + 1, generator=g))
    ref = orc.encode_rows(rays, t)                                       # [N*S, 123] fp32
    if source == "standalone_kernel":
        img = mlp_tc.encode_img(rays.cuda(), t.cuda())
    else:
        img = _chain_encoder_image(rays.cuda().contiguous(), t.cuda().contiguous())
    ipe, dirs = _decode_enc_image(img, N * S)
    for got, want in ((ipe, ref[:, :96]), (dirs[:, :27], ref[:, 96:])):
        err = (got - want).abs()
        bad = err > 2.0 ** -8 * want.abs() + 3e-5
        if bad.any():
            idx = bad.nonzero()[:8]
            detail = [(int(r), int(c), float(want[r, c]), float(got[r, c])) for r, c in idx]
            raise AssertionError(f"{int(bad.sum())} of {bad.numel()} features off by more than one bf16 step: (row, feature, want, got) {detail}")
    assert (dirs[:, 27:] == 0).all()


# ---------------------------------------------------------------------------------------------
# bf16 mode at the BENCHED configurations (BASELINE.json configs[1] and configs[2]), full size
# ---------------------------------------------------------------------------------------------
def _report(tag, got, ref, tol):
    err = (got - ref).abs()
    print(f"  {tag:18s} max {err.max().item():.3e}  mean {err.mean().item():.3e}  frac>{tol:g}: {(err > tol).float().mean().item():.2e}")
    return err


def test_bf16_cfg2_full_size_vs_oracle():
    """BASELINE.json configs[1] as bench.py times it: config_blender_mipnerf, 4096 rays, 128 + 128 samples, train mode, bf16
    tensor-core MLP -- against the fp32 oracle on the same weights / rays / random draws.  north_star's bf16 budget:
    sample depths (the FINE fence-posts produced from the bf16 coarse pass), rendered rgb and depth within 5e-3 max-abs;
    the loss within 5e-3; parameter-gradient direction against fp32 autograd."""
    from oracle import ddnerf_oracle as orc
    from ddnerf_b200.config import preset
    from ddnerf_b200.models import models as M
    from ddnerf_b200.rays import synth_rays
    cfg, kind = preset("config_blender_mipnerf", num_coarse=128, num_fine=128)
    N, s0, s1 = 4096, 128, 128
    ro, rd, rad, near, far = synth_rays(kind, N, seed=3)
    g = torch.Generator().manual_seed(0)
    target = torch.rand(N, 3, generator=g)
    rnd = dict(t_rand=torch.rand(N, s0 + 1, generator=g), noise0=torch.randn(N, s0, generator=g),
               u_rand=torch.rand(N, s1 + 1, generator=g), noise1=torch.randn(N, s1, generator=g))
    pc = orc.init_mlp_params(False, seed=7)
    tp = cfg.train_params
    ocfg = orc.PathConfig(model=cfg.nerf.type, near=near, far=far, num_coarse=s0, num_fine=s1, perturb=True,
                          noise_std=cfg.nerf.train.radiance_field_noise_std, blender=True, pdf_padding=tp.pdf_padding,
                          gaussian_smooth_factor=tp.gaussian_smooth_factor, dist_reg_coeficient=tp.dist_reg_coeficient,
                          loss_coeficients=tp.loss_coeficients, dp_coeficient=tp.dp_coeficient)
    rays = orc.pack_rays(ro, rd, rad, near, far)
    loss_ref, out_ref, gc_ref, _ = orc.train_step(ocfg, pc, None, rays, target, rnd)

    dev = torch.device("cuda:0")
    model = M.GeneralMipNerfModel(cfg)
    model.coarse.load_state_dict(pc)
    model.coarse.mlp_mode = "bf16"
    model.to(dev)
    model.keep_t_vals = True
    model.randoms = {k: v.to(dev) for k, v in rnd.items()}
    tgt = target.to(dev)
    out = model.run_iter(ro.to(dev), rd.to(dev), rad.to(dev), mode="train", rgb_target=tgt)
    loss = sum(tp.loss_coeficients[j] * torch.nn.functional.mse_loss(out[j]["rgb"], tgt) for j in range(2))
    loss.backward()
    torch.cuda.synchronize()
    print("cfg2 bf16 vs fp32 oracle (4096 rays x 128+128):")
    for j in range(2):
        t_got = torch.cat(model.last_t_vals[j]).cpu()
        e_t = _report(f"pass {j} t_vals", t_got, out_ref[j]["t_vals"], 5e-3)
        e_rgb = _report(f"pass {j} rgb", out[j]["rgb"].detach().cpu(), out_ref[j]["rgb"], 5e-3)
        e_dep = _report(f"pass {j} depth", out[j]["depth"].detach().cpu(), out_ref[j]["depth"], 5e-3)
        _report(f"pass {j} weights", out[j]["weights"].detach().cpu(), out_ref[j]["weights"], 5e-3)
        assert e_t.max().item() < 5e-3 and e_rgb.max().item() < 5e-3 and e_dep.max().item() < 5e-3
    assert abs(loss.item() - loss_ref.item()) < 5e-3
    flat_got = torch.cat([p.grad.reshape(-1) for p in model.coarse.parameters()]).cpu().double()
    flat_ref = torch.cat([gc_ref[k].reshape(-1) for k, _ in model.coarse.named_parameters()]).double()
    cos = torch.nn.functional.cosine_similarity(flat_got, flat_ref, dim=0).item()
    print(f"  whole-gradient cosine vs fp32 autograd: {cos:.5f}")
    assert cos > 0.97


def test_bf16_cfg3_ff_validation_slab_vs_oracle():
    """BASELINE.json configs[2] as bench.py's render leg runs it: config_ff (DDNeRF, forward-facing NDC rays), validation mode
    exactly as render_video.py:36-48 sets it up -- deterministic sampling, pdf_padding off, gaussian_smooth_factor =
    final_smooth (1.1), noise std 1.0 (injected) -- on a 17-row slab of the 1008 x 756 frame (17,136 rays, one chunk of
    16,384 + a ragged remainder), bf16 MLP, no_grad.  rgb / depth / disparity inputs and BOTH passes' sample depths within
    5e-3 of the fp32 oracle (t in [0, 1])."""
    from oracle import ddnerf_oracle as orc
    from ddnerf_b200.config import preset
    from ddnerf_b200.models import models as M
    from ddnerf_b200.rays import full_frame_rays
    cfg, kind = preset("config_ff")
    cfg.train_params.pdf_padding = False
    cfg.train_params.gaussian_smooth_factor = cfg.train_params.final_smooth
    ro, rd, rad, near, far = full_frame_rays(kind)
    r0, rows = 370, 17
    ro, rd, rad = (t[r0:r0 + rows].contiguous() for t in (ro, rd, rad))
    N = rows * ro.shape[1]
    g = torch.Generator().manual_seed(2)
    rnd = dict(noise0=torch.randn(N, 16, generator=g), noise1=torch.randn(N, 16, generator=g))
    pc, pf = orc.init_mlp_params(True, seed=7), orc.init_mlp_params(False, seed=8)
    tp = cfg.train_params
    ocfg = orc.PathConfig(model="DDNerfModel", near=near, far=far, num_coarse=16, num_fine=16, perturb=False, noise_std=1.0,
                          blender=False, pdf_padding=False, gaussian_smooth_factor=tp.final_smooth,
                          dist_reg_coeficient=tp.dist_reg_coeficient)
    with torch.no_grad():
        ref = orc.predict_dd(ocfg, pc, pf, orc.pack_rays(ro, rd, rad, near, far), rnd)

    dev = torch.device("cuda:0")
    model = M.DDNerfModel(cfg)
    model.coarse.load_state_dict(pc)
    model.fine.load_state_dict(pf)
    for net in (model.coarse, model.fine):
        net.mlp_mode = "bf16"
    model.to(dev)
    model.eval()
    model.record_distributions = False
    model.keep_t_vals = True
    model.randoms = {k: v.to(dev) for k, v in rnd.items()}
    with torch.no_grad():
        out = model.run_iter(ro.to(dev), rd.to(dev), rad.to(dev), mode="validation")
    torch.cuda.synchronize()
    assert out[1]["rgb"].shape == (rows, ro.shape[1], 3) and out[1]["disp"].shape == (rows, ro.shape[1])
    print(f"cfg3 bf16 validation slab ({N} rays x 16+16) vs fp32 oracle:")
    for j in range(2):
        t_got = torch.cat(model.last_t_vals[j]).cpu()
        e_t = _report(f"pass {j} t_vals", t_got, ref[j]["t_vals"], 5e-3)
        e_rgb = _report(f"pass {j} rgb", out[j]["rgb"].reshape(-1, 3).cpu(), ref[j]["rgb"], 5e-3)
        e_dep = _report(f"pass {j} depth", out[j]["depth"].reshape(-1).cpu(), ref[j]["depth"], 5e-3)
        e_acc = _report(f"pass {j} acc", out[j]["acc"].reshape(-1).cpu(), ref[j]["acc"], 5e-3)
        assert e_t.max().item() < 5e-3 and e_rgb.max().item() < 5e-3 and e_dep.max().item() < 5e-3 and e_acc.max().item() < 5e-3
    e_cd = _report("corrected disp", out[0]["corrected_disp_map"].reshape(-1).cpu(), ref[0]["corrected_disp_map"], 5e-3)
    rel = (e_cd / ref[0]["corrected_disp_map"].abs().clamp(min=1e-6)).max().item()
    assert rel < 5e-3


@pytest.mark.parametrize("N,S,depth_head", [(3, 128, False), (5, 256, True), (7, 128, False), (1, 16, True), (1024, 32, True),
                                            (2048, 128, False)])
def test_pair_kernels_bit_identical_to_single_cta(N, S, depth_head):
    """The CTA-pair chain kernels (cluster of 2, tcgen05 cta_group::2: M = 256 MMAs, weight chunks split across the
    pair, tiled TMA signalling the leader's barrier; the default) against the single-CTA kernels on the same inputs:
    forward outputs (training, inference with the in-kernel encoder, image-fed), saved activations, ReLU masks,
    encoded images and the dX chain's dZ images must be BIT-identical -- same K order, same epilogues.  Row counts
    that leave the last 512-row unit of a pair half or three quarters empty are included."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import check_pair
    from oracle import ddnerf_oracle as orc
    from ddnerf_b200 import _lib, mlp_tc
    from ddnerf_b200.models import base_architectures as BA
    from ddnerf_b200.rays import synth_rays
    lib = _lib.load()
    torch.manual_seed(N)
    ro, rd, rad, near, far = synth_rays("blender", N, seed=1)
    rays = orc.pack_rays(ro, rd, rad, near, far).cuda()
    t_vals = orc.sample_first_cycle(rays[:, 7:8].cpu(), rays[:, 8:9].cpu(), S).cuda()
    net = (BA.DepthMipNeRFModel if depth_head else BA.MipNeRFModel)(
        hidden_size=256, max_ipe_deg=16, num_encoding_fn_dir=4, include_input_xyz=False, include_input_dir=True)
    net.to("cuda")
    C = 6 if depth_head else 4
    st = mlp_tc._state(net)
    st.refresh()
    gout = torch.randn(N * S, C, device="cuda")
    prev = lib.ddnerf_mlp_tc_set_pair_mode(0)
    try:
        a = check_pair.run(lib, 0, C, rays, t_vals, st, gout)
        b = check_pair.run(lib, 1, C, rays, t_vals, st, gout)
    finally:
        lib.ddnerf_mlp_tc_set_pair_mode(prev)
    assert torch.isfinite(b["out"]).all()
    for k in a:
        assert torch.equal(a[k], b[k]), f"{k} differs between the single-CTA and the CTA-pair kernels"


@pytest.mark.parametrize("env", [{"DDNERF_TC_PAIR_SHARE": "1"}, {"DDNERF_TC_PSLOTS": "4"}])
def test_pair_kernel_schedule_knobs_keep_results(env):
    """The documented schedule knobs of the CTA-pair kernels -- weight stages shared by the two super-tiles (half a layer
    apart) instead of fetched per super-tile, and a shallower ring -- change the order of loads and waits only: outputs
    and saved images stay bit-identical to the single-CTA kernels (tools/check_pair.py in a fresh process, because the
    kernel programs are built once per process)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "check_pair.py")], env={**os.environ, **env}, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and "check_pair: OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_pair_dx_chain_on_a_few_ctas():
    """ddnerf_mlp_tc_backward_dx with max_ctas (the knob that lets a caller share the SMs with another kernel): 2, 5 and 37
    CTAs (the pair kernels round down to whole pairs) give the same dZ images as the full grid."""
    from oracle import ddnerf_oracle as orc
    from ddnerf_b200 import _lib, mlp_tc
    from ddnerf_b200.models import base_architectures as BA
    from ddnerf_b200.ops import _p, _stream
    lib = _lib.load()
    torch.manual_seed(3)
    rows, C = 256 * 21, 4
    net = BA.MipNeRFModel(hidden_size=256, max_ipe_deg=16, num_encoding_fn_dir=4, include_input_xyz=False, include_input_dir=True).to("cuda")
    st = mlp_tc._state(net)
    st.refresh()
    mask = torch.randint(0, 255, (lib.ddnerf_mlp_tc_mask_save_bytes(rows),), device="cuda", dtype=torch.uint8)
    gout = torch.randn(rows, C, device="cuda")
    outs = []
    for ctas in (0, 2, 5, 37):
        dz = torch.zeros(lib.ddnerf_mlp_tc_act_save_bytes(rows), device="cuda", dtype=torch.uint8)
        _lib.check(lib.ddnerf_mlp_tc_backward_dx(_p(st.wimg), _p(st.bias), _p(gout), rows, C, _p(mask), _p(dz), ctas, _stream()), "dx")
        torch.cuda.synchronize()
        outs.append(dz)
    for dz in outs[1:]:
        assert torch.equal(dz, outs[0])

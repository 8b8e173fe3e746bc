"""GPU suite: the CUDA path (through the C ABI) against the golden vectors made from the reference
and against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): fp32 path 1e-3 max-abs on sample depths and rendered
rgb/depth -- the per-kernel checks below are much tighter (1e-5..1e-4) because each kernel
follows the reference's fp32 arithmetic op by op; search indices are bit-exact given identical
CDFs.
"""
import numpy as np
import pytest
import torch

from oracle import ddnerf_oracle as orc
from tests.conftest import load_golden
from tests.test_oracle_golden import E2E, e2e_setup

pytestmark = pytest.mark.gpu

DEV = "cuda"


def cu(t):
    return t.to(DEV) if isinstance(t, torch.Tensor) else t


def close(a, b, rtol=1e-5, atol=1e-5):
    torch.testing.assert_close(a.detach().cpu(), b.detach().cpu(), rtol=rtol, atol=atol, equal_nan=True)


@pytest.fixture(scope="module")
def ops():
    from ddnerf_b200 import ops as _ops
    return _ops


def test_library_on_device(ops):
    from ddnerf_b200 import _lib
    lib = _lib.load()
    assert lib.ddnerf_version() == 1
    assert lib.ddnerf_device_is_sm100() == 1, "tests must run on a B200 (sm_100)"


def test_cpu_tensor_raises(ops):
    with pytest.raises(RuntimeError):
        ops.encode(torch.zeros(2, 12), torch.zeros(2, 5))


# ---------------------------------------------------------------------------------------------
# K3 samplers
# ---------------------------------------------------------------------------------------------
def test_first_cycle(ops):
    g = load_golden("first_cycle")
    near, far = cu(g["near"]), cu(g["far"])
    close(ops.sample_first_cycle(near, far, 16), g["fc_det_t"].expand(24, 17), 1e-6, 1e-6)
    close(ops.sample_first_cycle(near, far, 32, False, cu(g["fc_jit_rand"])), g["fc_jit_t"], 1e-6, 1e-6)
    close(ops.sample_first_cycle(near, far, 8, True, cu(g["fc_lind_rand"])), g["fc_lind_t"], 1e-6, 1e-6)
    # strided near/far taken straight from a packed [N,12] ray tensor
    rays = torch.zeros(24, 12, device=DEV)
    rays[:, 7], rays[:, 8] = 2.0, 6.0
    close(ops.sample_first_cycle(rays[:, 7:8], rays[:, 8:9], 16), g["fc_det_t"].expand(24, 17), 1e-6, 1e-6)


@pytest.mark.parametrize("S", [16, 32, 48])
@pytest.mark.parametrize("wname", ["uniform", "peaked"])
@pytest.mark.parametrize("pad", [True, False])
@pytest.mark.parametrize("det", [True, False])
def test_resamplers_golden(ops, S, wname, pad, det):
    g = load_golden(f"resample_S{S}")
    key = f"{wname}_pad{int(pad)}_det{int(det)}"
    bins, w = cu(g["bins"]), cu(g["w_" + wname])
    n = g["mip_" + key].shape[1]
    rand = None if det else cu(g["mip_" + key + "_rand"])
    close(ops.sample_pdf(bins, w, n, pad, rand), g["mip_" + key], 1e-5, 2e-5)
    rand = None if det else cu(g["dd_" + key + "_rand"])
    s = ops.sample_pdf_mu_sigma(bins, w, cu(g["mus"]), cu(g["sigmas"]), cu(g["pin"]), cu(g["lt"]), n, pad,
                                float(g["near_cfg"]), float(g["far_cfg"]), rand)
    # erfinv near its singularities amplifies 1-ulp cdf differences; depths are in [2,6]
    close(s, g["dd_" + key], 1e-4, 2e-4)


def test_resampler_one_cell(ops):
    g = load_golden("resample_onecell")
    s = ops.sample_pdf_mu_sigma(cu(g["bins"]), cu(g["w"]), cu(g["mus"]), cu(g["sigmas"]), cu(g["pin"]), cu(g["lt"]), 9,
                                True, float(g["near_cfg"]), float(g["far_cfg"]))
    close(s, g["dd"], 1e-4, 2e-4)


@pytest.mark.parametrize("N,S,n", [(257, 16, 17), (1000, 64, 65), (513, 128, 129), (64, 1, 9)])
def test_find_interval_bit_exact(ops, N, S, n):
    """Same CDF in, same interval index out as the reference's mask search (oracle restatement)."""
    gen = torch.Generator().manual_seed(N + S)
    w = torch.rand(N, S, generator=gen) ** 3
    w[:, S // 3: S // 2] = 0.0                      # empty space -> flat (tied) CDF segments
    cdf = orc.resampling_cdf(w, pdf_padding=False) if S > 1 else torch.tensor([[0.0, 1.0]]).expand(N, 2).contiguous()
    u = torch.rand(N, n, generator=gen)
    u[:, 0], u[:, -1] = 0.0, 1.0
    u[:, 1] = cdf[:, min(1, S)]                      # exactly on an edge
    j_ref, _ = orc.find_interval(cdf, u)
    idx = ops.find_interval(cu(cdf), cu(u))
    assert torch.equal(idx.cpu().long(), j_ref)


def test_resampler_indices_match_oracle(ops):
    g = load_golden("resample_S32")
    for wname in ("uniform", "peaked"):
        w = g["w_" + wname]
        n = 33
        s_ref, j_ref = orc.sample_pdf(g["bins"], w, n, True, None)
        s, j = ops.sample_pdf(cu(g["bins"]), cu(w), n, True, None, return_idx=True)
        # indices may differ only where u sits within float rounding of a CDF edge (the CUDA CDF is a
        # warp scan in fp32, the CPU one accumulates in double); samples must agree regardless
        frac = (j.cpu().long() != j_ref).float().mean().item()
        assert frac < 0.01
        close(s, s_ref, 1e-5, 2e-5)


def _resample_inputs(N, S, seed, peaked):
    g = torch.Generator().manual_seed(seed)
    bins = torch.sort(torch.rand(N, S + 1, generator=g) * 4 + 2, dim=-1)[0]
    bins[:, 0], bins[:, -1] = 2.0, 6.0
    w = torch.rand(N, S, generator=g) ** (6 if peaked else 1)
    if S >= 6:
        w[:, S // 3: S // 2] = 0.0                   # empty space
    mus = torch.rand(N, S, generator=g)
    sig = torch.rand(N, S, generator=g) * 0.5 + 1e-3
    lt = orc.normal_cdf((0 - mus) / sig)
    pin = orc.normal_cdf((1 - mus) / sig) - lt
    return g, bins, w, mus, sig, lt, pin


# every lane-group shape of the fast path (S <= 256: G x C = 4x1 .. 32x8, K = 5 / 9), ragged S and N, and the
# generic kernels behind them (S = 300; n > 9 G)
@pytest.mark.parametrize("S,n", [(2, 9), (3, 4), (7, 8), (16, 17), (31, 32), (33, 34), (64, 65), (100, 101), (128, 129),
                                 (128, 40), (200, 201), (256, 257), (300, 301), (16, 200)])
@pytest.mark.parametrize("det", [True, False])
def test_resamplers_shapes_vs_oracle(ops, S, n, det):
    N = 37
    g, bins, w, mus, sig, lt, pin = _resample_inputs(N, S, 100 + S + n, peaked=(S % 2 == 0))
    rand = None if det else torch.rand(N, n, generator=g)
    for pad in (True, False):
        s_ref, _ = orc.sample_pdf(bins, w, n, pad, rand)
        close(ops.sample_pdf(cu(bins), cu(w), n, pad, cu(rand)), s_ref, 1e-5, 2e-5)
        d_ref, _ = orc.sample_pdf_with_mu_sigma(bins, w, mus, sig, pin, lt, n, pad, 2.0, 6.0, rand)
        d = ops.sample_pdf_mu_sigma(cu(bins), cu(w), cu(mus), cu(sig), cu(pin), cu(lt), n, pad, 2.0, 6.0, cu(rand))
        assert (d[:, 1:] >= d[:, :-1]).all()
        # erfinv near its singularities amplifies 1-ulp cdf differences (see test_resamplers_golden): all but a
        # handful of samples within 3e-4, none off by more than 5e-2 (depths are in [2,6])
        err = (d.cpu() - d_ref).abs()
        assert (err > 3e-4 + 1e-4 * d_ref.abs()).float().mean().item() < 2e-3 and err.max().item() < 5e-2


def _error_distributions(got, ref32, exact):
    """Sorted |error| of the CUDA result and of the reference's own fp32 evaluation, both against exact (fp64) arithmetic
    on the same fp32 inputs."""
    e_got = np.sort((got.double() - exact).abs().flatten().numpy())
    e_ref = np.sort((ref32.double() - exact).abs().flatten().numpy())
    return e_got, e_ref


@pytest.mark.parametrize("S,n", [(16, 17), (32, 33), (64, 65), (128, 129)])
@pytest.mark.parametrize("det", [True, False])
def test_dd_resampler_error_vs_fp64_yardstick(ops, S, n, det):
    """sample_pdf_with_mu_sigma (samplers.py:124-215) puts Phi^-1 behind a CDF interpolation: near z -> 0 / 0.999 and where
    (c1 - c0) is tiny, one ulp of the CDF moves a sample by far more than one ulp, in the reference's own fp32 evaluation
    as much as in the kernel.  Yardstick: the same formulas in float64 on the same fp32 inputs.  The kernel's error
    DISTRIBUTION must be no worse than 3x the distribution of the reference's fp32 errors at every quantile up to 99.99 %
    (floor: three ulps of the depth range).  The last 0.01 % are a handful of samples sitting on a singularity, where which
    of the two evaluations gets the unlucky rounding is a coin toss (measured maxima over the cases of this test: 2.6e-2 for
    the reference's fp32 evaluation, 2.6e-2 for the kernel, in different cases): there the FREQUENCY of outliers is held
    -- at each of the thresholds 1e-4 / 1e-3 / 1e-2 no more than 3x the reference's own count + 3 samples (of 70-530 k)
    -- and nothing may be off by more than 5e-2.  4096 rays per case (both smoothing branches, flat and peaked weights
    with empty space, sigmas down to 1e-3)."""
    N = 4096
    worst = 0.0
    for peaked in (False, True):
        g, bins, w, mus, sig, lt, pin = _resample_inputs(N, S, 900 + S + int(peaked), peaked=peaked)
        rand = None if det else torch.rand(N, n, generator=g)
        for pad in (True, False):
            args32 = (bins, w, mus, sig, pin, lt)
            ref32, _ = orc.sample_pdf_with_mu_sigma(*args32, n, pad, 2.0, 6.0, rand)
            exact, _ = orc.sample_pdf_with_mu_sigma(*[a.double() for a in args32], n, pad, 2.0, 6.0,
                                                    None if rand is None else rand.double())
            got = ops.sample_pdf_mu_sigma(*[cu(a) for a in args32], n, pad, 2.0, 6.0, cu(rand)).cpu()
            assert (got[:, 1:] >= got[:, :-1]).all()
            e_got, e_ref = _error_distributions(got, ref32, exact)
            floor = 3 * 4.8e-7                                        # 3 ulp at depth 6
            body = int(0.9999 * len(e_got))
            ratio = float(np.max(e_got[:body] / (3.0 * e_ref[:body] + floor)))
            for thr in (1e-4, 1e-3, 1e-2):
                n_got, n_ref = int((e_got > thr).sum()), int((e_ref > thr).sum())
                assert n_got <= 3 * n_ref + 3, (S, n, det, peaked, pad, thr, n_got, n_ref)
            assert e_got[-1] < 5e-2
            worst = max(worst, ratio)
            q = [0.5, 0.99, 0.999, 1.0]
            idx = [min(len(e_got) - 1, int(x * len(e_got))) for x in q]
            print(f"S={S} n={n} det={det} peaked={peaked} pad={pad}: quantiles {q} kernel {[f'{e_got[i]:.2e}' for i in idx]} "
                  f"reference fp32 {[f'{e_ref[i]:.2e}' for i in idx]}  worst ratio to 3x yardstick {ratio:.2f}")
            assert ratio <= 1.0, (S, n, det, peaked, pad, ratio)
    print(f"worst ratio {worst:.2f}")


@pytest.mark.parametrize("S,n", [(32, 33), (128, 129)])
def test_mip_resampler_error_vs_fp64_yardstick(ops, S, n):
    """The same yardstick for sample_pdf (samplers.py:64-121)."""
    N = 4096
    for peaked in (False, True):
        g, bins, w, *_ = _resample_inputs(N, S, 700 + S + int(peaked), peaked=peaked)
        rand = torch.rand(N, n, generator=g)
        for pad in (True, False):
            ref32, _ = orc.sample_pdf(bins, w, n, pad, rand)
            exact, _ = orc.sample_pdf(bins.double(), w.double(), n, pad, rand.double())
            got = ops.sample_pdf(cu(bins), cu(w), n, pad, cu(rand)).cpu()
            e_got, e_ref = _error_distributions(got, ref32, exact)
            body = int(0.9999 * len(e_got))
            ratio = max(float(np.max(e_got[:body] / (3.0 * e_ref[:body] + 3 * 4.8e-7))), float(e_got[-1] / (10.0 * e_ref[-1] + 3 * 4.8e-7)))
            assert int((e_got > 1e-4).sum()) <= 3 * int((e_ref > 1e-4).sum()) + 3
            print(f"mip S={S} peaked={peaked} pad={pad}: max kernel {e_got[-1]:.2e} reference fp32 {e_ref[-1]:.2e} ratio {ratio:.2f}")
            assert ratio <= 1.0


def test_dd_resampler_sorts_when_bins_leave_cfg_range(ops):
    """samplers.py:210-213: endpoints pinned to cfg near/far, then sorted -- exercised with a cfg range INSIDE the bins."""
    N, S, n = 19, 32, 33
    g, bins, w, mus, sig, lt, pin = _resample_inputs(N, S, 5, peaked=False)
    rand = torch.rand(N, n, generator=g)
    d_ref, _ = orc.sample_pdf_with_mu_sigma(bins, w, mus, sig, pin, lt, n, True, 3.0, 5.0, rand)
    d = ops.sample_pdf_mu_sigma(cu(bins), cu(w), cu(mus), cu(sig), cu(pin), cu(lt), n, True, 3.0, 5.0, cu(rand))
    assert (d[:, 1:] >= d[:, :-1]).all()
    close(d, d_ref, 1e-4, 3e-4)


@pytest.mark.parametrize("S", [1, 5, 16, 32, 33, 100, 128, 129, 257, 300])
def test_first_cycle_shapes_vs_oracle(ops, S):
    N = 45
    g = torch.Generator().manual_seed(S)
    near, far = torch.rand(N, 1, generator=g) + 1.5, torch.rand(N, 1, generator=g) + 5.0
    rnd = torch.rand(N, S + 1, generator=g)
    for lind in (False, True):
        close(ops.sample_first_cycle(cu(near), cu(far), S, lind, cu(rnd)), orc.sample_first_cycle(near, far, S, lind, rnd), 1e-6, 1e-6)
        close(ops.sample_first_cycle(cu(near), cu(far), S, lind), orc.sample_first_cycle(near, far, S, lind), 1e-6, 1e-6)


def test_first_cycle_large_and_unaligned(ops):
    """The flat kernel's eight-per-thread variant (>= 4 M fence-posts), per-ray near / far, and buffers that are not 32-byte
    aligned (scalar loads / stores instead of 16-byte ones)."""
    N, S = 33000, 128
    g = torch.Generator().manual_seed(77)
    near, far = torch.rand(N, 1, generator=g) + 1.5, torch.rand(N, 1, generator=g) + 5.0
    rnd = torch.rand(N, S + 1, generator=g)
    ref = orc.sample_first_cycle(near, far, S, False, rnd)
    close(ops.sample_first_cycle(cu(near), cu(far), S, False, cu(rnd)), ref, 1e-6, 1e-6)
    pad = torch.zeros(N * (S + 1) + 3, device=DEV)
    pad[1:1 + N * (S + 1)] = cu(rnd).flatten()
    rnd_odd = pad[1:1 + N * (S + 1)].view(N, S + 1)                    # data pointer 4 bytes past an aligned address
    assert rnd_odd.data_ptr() % 16 == 4
    close(ops.sample_first_cycle(cu(near), cu(far), S, False, rnd_odd), ref, 1e-6, 1e-6)
    M = 1000                                                           # small, unaligned: the four-per-thread variant
    rnd_odd = pad[1:1 + M * (S + 1)].view(M, S + 1)
    close(ops.sample_first_cycle(cu(near[:M]), cu(far[:M]), S, False, rnd_odd),
          orc.sample_first_cycle(near[:M], far[:M], S, False, rnd[:M]), 1e-6, 1e-6)


# ---------------------------------------------------------------------------------------------
# f1 ray generation
# ---------------------------------------------------------------------------------------------
def test_ray_bundle_golden(ops):
    """csrc/raygen.cu against what the reference's get_ray_bundle / ndc_mipnerf_rays returned (tests/golden)."""
    from ddnerf_b200 import rays as R
    g = load_golden("ray_bundle")
    H, W, focal = int(g["p_H"]), int(g["p_W"]), float(g["p_focal"])
    ro, rd, rad = R.ray_bundle_cuda(H, W, focal, g["p_pose"])
    for got, key in ((ro, "p_ro"), (rd, "p_rd"), (rad, "p_rad")):
        close(got, g[key].reshape(got.shape), 1e-6, 1e-7)
    H, W, focal = int(g["n_H"]), int(g["n_W"]), float(g["n_focal"])
    o, d, r = R.ray_bundle_cuda(H, W, focal, g["n_pose"], ndc_near=1)
    for got, key in ((o, "n_ro"), (d, "n_rd"), (r, "n_rad")):
        close(got, g[key].reshape(got.shape), 2e-6, 1e-6)
    # a pose on the device routes the reference-named entry point to the kernel
    from ddnerf_b200.general_utils import nerf_helpers as H_
    ro2, rd2, rad2 = H_.get_ray_bundle(int(g["p_H"]), int(g["p_W"]), float(g["p_focal"]), cu(g["p_pose"]))
    close(ro2, ro, 0, 0); close(rd2, rd, 0, 0)


@pytest.mark.parametrize("kind", ["blender", "ff", "360"])
def test_ray_bundle_full_frames_and_row_split(ops, kind):
    """BASELINE.json frame sizes: the kernel against the host restatement, and rows [lo,hi) == the same rows of the frame."""
    from ddnerf_b200 import rays as R
    H, W, focal, c2w, near, far, ndc = R.frame(kind)
    ro_ref, rd_ref, rad_ref, _, _ = R.full_frame_rays(kind)
    ro, rd, rad = R.ray_bundle_cuda(H, W, focal, c2w, ndc_near=1 if ndc else None)
    close(ro, ro_ref, 2e-6, 1e-6); close(rd, rd_ref, 2e-6, 1e-6); close(rad, rad_ref, 2e-6, 1e-9)
    lo, hi = H // 3, H - 5
    ro_p, rd_p, rad_p = R.ray_bundle_cuda(H, W, focal, c2w, rows=(lo, hi), ndc_near=1 if ndc else None)
    assert torch.equal(ro_p, ro[lo:hi]) and torch.equal(rd_p, rd[lo:hi]) and torch.equal(rad_p, rad[lo:hi])
    ro_t, _, rad_t = R.ray_bundle_cuda(H, W, focal, c2w, rows=(H - 1, H), ndc_near=1 if ndc else None)
    assert torch.equal(rad_t, rad[H - 1:]) and torch.equal(ro_t, ro[H - 1:])


# ---------------------------------------------------------------------------------------------
# f4 frame post-processing and the pose -> 8-bit frame loop
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("H,W", [(3, 5), (64, 48), (756, 1008)])
def test_frame_to_u8_vs_oracle(ops, H, W):
    from ddnerf_b200.render import frame_to_u8
    g = torch.Generator().manual_seed(H)
    rgb = torch.rand(H, W, 3, generator=g) * 1.002 - 0.001          # the renderer's colour range
    rgb[0, 0] = torch.tensor([-0.001, 1.001, 0.5])
    disp = 1.0 / (torch.rand(H, W, generator=g) * 5 + 0.2)
    rgb8, disp8, video = frame_to_u8(cu(rgb), cu(disp), want_video=True)
    d_ref = orc.cast_to_disparity_image(disp)
    assert np.array_equal(rgb8.cpu().numpy(), orc.cast_to_image(rgb))
    assert np.array_equal(disp8.cpu().numpy(), d_ref)
    assert np.array_equal(video.cpu().numpy(), orc.video_frame(rgb, d_ref))
    assert disp8.min().item() == 0 and disp8.max().item() == 255


def test_frame_to_u8_reference_golden(ops):
    """csrc/frame.cu against what the REFERENCE's cast_to_image / cast_to_disparity_image (visualization.py:11-27) and
    its video-frame assembly (render_video.py:98-103) returned for the same float frame (tests/golden/frame_post.npz)."""
    from ddnerf_b200.render import frame_to_u8
    g = load_golden("frame_post")
    rgb8, disp8, video = frame_to_u8(cu(g["rgb"]), cu(g["disp"]), want_video=True)
    assert np.array_equal(np.moveaxis(rgb8.cpu().numpy(), -1, 0), g["rgb8_chw"].numpy())
    assert np.array_equal(disp8.cpu().numpy()[None], g["disp8"].numpy())
    assert np.array_equal(video.cpu().numpy(), g["video_bgr"].numpy())


def test_frame_renderer_pose_to_u8(ops):
    """pose -> rays -> model -> 8-bit frame as one replayed CUDA graph == the same steps done by hand from host rays."""
    from ddnerf_b200 import rays as R
    from ddnerf_b200.config import preset
    from ddnerf_b200.models import models as M
    from ddnerf_b200.render import FrameRenderer
    cfg, _ = preset("config_ff")
    cfg.train_params.pdf_padding = False
    cfg.nerf.validation.radiance_field_noise_std = 0.0               # deterministic frames
    H, W, focal = 24, 40, 31.0
    torch.manual_seed(0)
    model = M.DDNerfModel(cfg)
    model.to(DEV)
    model.eval()
    for net in (model.coarse, model.fine):
        net.mlp_mode = "fp32"
    fr = FrameRenderer(model, H, W, focal, ndc_near=1, want_video=True, use_graph=True)
    poses = []
    for k in range(4):
        p = torch.eye(4)
        p[:3, 3] = torch.tensor([0.05 * k, -0.03 * k, 0.01])
        poses.append(p)
    for k, p in enumerate(poses):                                    # calls 0,1 eager, 2 captures, 3 replays
        rgb8, disp8, video = (t.clone() for t in fr.render(p))
        ro, rd, _ = R.get_ray_bundle(H, W, focal, p)
        o, d, r = R.ndc_mipnerf_rays(H, W, focal, ro, rd, near=1)
        with torch.no_grad():
            out = model.run_iter(cu(o), cu(d), cu(r), mode="validation")
        rgb, disp = out[1]["rgb"].cpu(), out[1]["disp"].cpu()
        d_ref = orc.cast_to_disparity_image(disp)
        # the device rays differ from the host restatement in the last ulp, so allow one grey level on a few pixels
        assert (np.abs(rgb8.numpy().astype(int) - orc.cast_to_image(rgb).astype(int)) > 1).mean() == 0, k
        assert (np.abs(disp8.numpy().astype(int) - d_ref.astype(int)) > 1).mean() < 0.01, k
        assert np.array_equal(video.numpy()[:, :W, ::-1], rgb8.numpy()) and np.array_equal(video.numpy()[:, W:, 0], disp8.numpy())
    assert fr._graph is not None


# ---------------------------------------------------------------------------------------------
# f3 device-resident training ray store
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ndc", [False, True])
def test_ray_store_matches_train_dataset(ops, ndc):
    """dataset.py:8-59: vstack of every image's rays + targets, batch = rows[idxs]."""
    from ddnerf_b200 import rays as R
    from ddnerf_b200.raystore import DeviceRayStore
    g = torch.Generator().manual_seed(3)
    n_img, H, W, focal = 3, 10, 14, 12.5
    poses = []
    for k in range(n_img):
        p = torch.eye(4) if ndc else R.pose_spherical(20.0 * k, -15.0, 3.0)
        p = p.clone()
        p[:3, 3] += torch.tensor([0.05 * k + 0.01, 0.02, 0.03])
        poses.append(p)
    poses = torch.stack(poses)
    images = torch.rand(n_img, H, W, 3, generator=g)
    ref = [[], [], [], []]
    for k in range(n_img):
        ro, rd, rad = R.get_ray_bundle(H, W, focal, poses[k])
        if ndc:
            ro, rd, rad = R.ndc_mipnerf_rays(H, W, focal, ro, rd, 1)
        for lst, t in zip(ref, (ro.reshape(-1, 3), rd.reshape(-1, 3), rad.reshape(-1, 1), images[k].reshape(-1, 3))):
            lst.append(t)
    ref = [torch.vstack(x) for x in ref]
    store = DeviceRayStore(poses, images, focal, ndc_rays=ndc)
    assert len(store) == n_img * H * W
    idxs = torch.randint(0, len(store), (257,), generator=g)
    idxs[0], idxs[1] = 0, len(store) - 1
    out = store.get_training_rays_for_next_iter(257, DEV, idxs=idxs)
    for got, want in zip(out, ref):
        close(got, want[idxs], 2e-6, 1e-6)
    assert torch.equal(out[3].cpu(), ref[3][idxs])                       # targets are copied, not recomputed
    # device-side draws: right shapes, every row is a row of the store
    ro, rd, rad, tgt = store.get_training_rays_for_next_iter(1000, DEV)
    assert ro.shape == (1000, 3) and rad.shape == (1000, 1) and tgt.shape == (1000, 3)
    full = store.get_training_rays_for_next_iter(len(store), DEV, idxs=torch.arange(len(store)))
    for i in range(50):
        assert ((full[1] == rd[i]).all(1) & (full[3] == tgt[i]).all(1) & (full[2] == rad[i]).all(1)).any()
    assert store.check_indices()
    store.get_training_rays_for_next_iter(2, DEV, idxs=torch.tensor([0, len(store)]))
    assert not store.check_indices()                                     # out-of-range index is flagged, not read
    # single-image mode: rows of ONE image
    s1 = DeviceRayStore(poses, images, focal, ndc_rays=ndc, single_image_mode=True)
    sub = torch.randint(0, H * W, (64,), generator=g)
    out1 = s1.get_training_rays_for_next_iter(64, DEV, idxs=sub, img_idx=2)
    close(out1[0], ref[0][2 * H * W + sub], 2e-6, 1e-6)
    assert torch.equal(out1[3].cpu(), ref[3][2 * H * W + sub])


# ---------------------------------------------------------------------------------------------
# K2 encoding
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["blender", "ff", "360"])
@pytest.mark.parametrize("shape", ["cone", "cylinder"])
def test_encoding(ops, kind, shape):
    g = load_golden(f"encoding_{kind}")
    N = g["ro"].shape[0]
    rays = orc.pack_rays(g["ro"], g["rd"], g["rad"], 0.0, 1.0)
    x = ops.encode(cu(rays), cu(g["t"]), shape)
    S = g["t"].shape[1] - 1
    assert x.shape == (N * S, 123)
    close(x[:, :96].reshape(N, S, 96), g[f"ipe_{shape}"], 1e-5, 1e-5)
    close(x[:, 96:].reshape(N, S, 27)[:, 0], g["dir_enc"], 1e-6, 1e-6)


# ---------------------------------------------------------------------------------------------
# K1 MLP (fp32 mode)
# ---------------------------------------------------------------------------------------------
def _module(depth, seed):
    from ddnerf_b200.models import base_architectures as arch
    net = (arch.DepthMipNeRFModel if depth else arch.MipNeRFModel)(
        hidden_size=256, max_ipe_deg=16, num_encoding_fn_dir=4, include_input_xyz=False, include_input_dir=True,
        use_viewdirs=True)
    net.load_state_dict(orc.init_mlp_params(depth, seed=seed))
    return net.to(DEV)


@pytest.mark.parametrize("depth", [False, True])
def test_mlp_f32_golden(ops, depth):
    g = load_golden("mlp_depth" if depth else "mlp_plain")
    net = _module(depth, 11 + int(depth))
    x = cu(g["x"]).requires_grad_(True)
    y = net(x)
    close(y, g["y"], 1e-4, 1e-5)
    y.backward(cu(g["gy"]))
    for k, p in net.named_parameters():
        if "g_" + k in g:
            close(p.grad, g["g_" + k], 1e-3, 1e-4)
    # dx against the oracle
    params = {k: v.requires_grad_(False) for k, v in orc.init_mlp_params(depth, seed=11 + int(depth)).items()}
    xr = g["x"].clone().requires_grad_(True)
    orc.mlp_forward(params, xr).backward(g["gy"])
    close(x.grad, xr.grad, 1e-3, 1e-4)


def test_mlp_f32_ragged_rows(ops):
    """rows not a multiple of any tile size; fused encode path == explicit feature path."""
    net = _module(True, 5)
    N, S = 37, 7
    g = torch.Generator().manual_seed(3)
    from ddnerf_b200.rays import synth_rays
    ro, rd, rad, near, far = synth_rays("blender", N, seed=5)
    rays = cu(orc.pack_rays(ro, rd, rad, near, far))
    t = cu(orc.sample_first_cycle(torch.full((N, 1), near), torch.full((N, 1), far), S, False,
                                  torch.rand(N, S + 1, generator=g)))
    y1 = net.forward_rays(rays, t, "cone")
    y2 = net(ops.encode(rays, t, "cone")).reshape(N, S, 6)
    close(y1, y2, 1e-6, 1e-6)
    params = orc.init_mlp_params(True, seed=5)
    close(y1, orc.run_network(params, rays.cpu(), t.cpu()), 1e-4, 2e-5)


# ---------------------------------------------------------------------------------------------
# K4 compositing
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,std,white,use_mus,blender", [
    ("blender", 0.0, False, False, True), ("blender_noise_white", 1.0, True, False, True),
    ("blender_mus", 1.0, False, True, True), ("real_mus", 0.5, False, True, False),
    ("nocfg", 0.0, False, False, False)])
def test_composite_golden(ops, tag, std, white, use_mus, blender):
    g = load_golden("render")
    raw = cu(g["raw"]).requires_grad_(True)
    mus = cu(g["mus"]).requires_grad_(True)
    noise = cu(g[f"{tag}_randn"]) if std > 0 else None
    outs = ops.composite(raw, cu(g["t"]), cu(g["rd"]), noise, std, mus if use_mus else None, white, blender, True)
    names = ("rgb_map", "disp", "acc", "weights", "depth", "cdisp", "rgb")
    loss = 0
    for nme, o in zip(names, outs):
        if o is None:
            assert f"{tag}_{nme}" not in g
            continue
        ref = g[f"{tag}_{nme}"]
        if nme in ("disp", "cdisp"):
            close(1.0 / o, 1.0 / ref, 1e-4, 1e-5)
        else:
            close(o, ref, 1e-4, 1e-6)
        if f"{tag}_ct_{nme}" in g:
            loss = loss + (o * cu(g[f"{tag}_ct_{nme}"])).sum()
    loss.backward()
    gr, gref = raw.grad.cpu(), g[f"{tag}_g_raw"]
    # rays with sum(weights) ~ 0 have disp ~ 1e10: compare gradients relative to each ray's scale
    scale = gref.nan_to_num(0.0).abs().amax(dim=(1, 2), keepdim=True).clamp(min=1e-6)
    finite = torch.isfinite(gref)
    assert torch.equal(torch.isfinite(gr), finite)
    err = ((gr - gref).abs() / scale)[finite].max().item()
    assert err < 2e-4, err
    if use_mus:
        gm, gmref = mus.grad.cpu(), g[f"{tag}_g_mus"]
        sc = gmref.nan_to_num(0.0).abs().amax(dim=1, keepdim=True).clamp(min=1e-6)
        fin = torch.isfinite(gmref)
        assert ((gm - gmref).abs() / sc)[fin].max().item() < 2e-4


@pytest.mark.parametrize("S", [1, 7, 16, 33, 64, 128, 200, 512])
def test_composite_shapes_vs_oracle(ops, S):
    N = 65
    g = torch.Generator().manual_seed(S)
    t = torch.sort(torch.rand(N, S + 1, generator=g) * 4 + 2, dim=-1)[0]
    raw6 = torch.randn(N, S, 6, generator=g) * 2
    rd = torch.randn(N, 3, generator=g)
    mus = torch.rand(N, S, generator=g)
    nz = torch.randn(N, S, generator=g)
    ref = orc.volume_render(raw6[..., :4], t, rd, nz * 0.7, True, True, mus)
    out = ops.composite(cu(raw6)[..., :4], cu(t), cu(rd), cu(nz), 0.7, cu(mus), True, True, True)   # strided raw view
    for o, r, nme in zip(out, ref, ("rgb_map", "disp", "acc", "weights", "depth", "cdisp", "rgb")):
        if nme in ("disp", "cdisp"):
            close(1.0 / o, 1.0 / r, 1e-4, 1e-5)
        else:
            close(o, r, 1e-4, 2e-6)


# ---------------------------------------------------------------------------------------------
# K4 + the DDNeRF depth-distribution head (coarse pass of DDNerfModel.predict, models.py:242-273) as one kernel
# ---------------------------------------------------------------------------------------------
def _dd_coarse_reference(raw6, t, rd, nz, std, white, blender, coef):
    """models.py:242-264 restated with the oracle's pieces (torch fp32 on the CPU, autograd for the backward)."""
    raw_mus, raw_sig = raw6[..., -2], raw6[..., -1]
    mus = torch.sigmoid(raw_mus)
    sigmas = torch.sigmoid(raw_sig) + 0.001
    sig_loss = (torch.abs(raw_sig) ** 2).sum() / raw_sig.shape[0]
    mus_loss = (torch.abs(raw_mus) ** 2).sum() / raw_mus.shape[0]
    regs = torch.stack([mus_loss, sig_loss, coef * mus_loss, coef * sig_loss])
    rgb_map, disp, acc, w, depth, cdisp, _ = orc.volume_render(raw6[..., :4], t, rd, nz * std if std > 0 else None, white,
                                                               blender, mus)
    return rgb_map, disp, acc, w, depth, cdisp, mus, sigmas, regs


@pytest.mark.parametrize("S", [1, 7, 16, 32, 33, 64, 128, 200])
@pytest.mark.parametrize("blender,white,std", [(True, False, 1.0), (False, True, 0.0)])
def test_composite_dd_vs_oracle(ops, S, blender, white, std):
    """ops.composite_dd forward and backward against the reference's op sequence: all nine outputs, and the cotangent of
    the [N,S,6] network output for a loss that touches every output (rgb, weights, depth, disparities, mus, sigmas, the
    four regulariser scalars) -- sigmoid' of the head and the 2 raw / N of the regularisers included."""
    N = 67
    g = torch.Generator().manual_seed(1000 + S)
    t = torch.sort(torch.rand(N, S + 1, generator=g) * 4 + 2, dim=-1)[0]
    raw6 = torch.randn(N, S, 6, generator=g) * 2
    rd = torch.randn(N, 3, generator=g)
    nz = torch.randn(N, S, generator=g)
    coef = 0.0625
    cts = [torch.randn(N, 3, generator=g), torch.randn(N, generator=g) * 0.01, torch.randn(N, generator=g),
           torch.randn(N, S, generator=g), torch.randn(N, generator=g), torch.randn(N, generator=g) * 0.01,
           torch.randn(N, S, generator=g), torch.randn(N, S, generator=g), torch.randn(4, generator=g)]
    a = raw6.clone().requires_grad_(True)
    ref = _dd_coarse_reference(a, t, rd, nz, std, white, blender, coef)
    sum((o * c).sum() for o, c in zip(ref, cts)).backward()
    b = cu(raw6).clone().requires_grad_(True)
    out = ops.composite_dd(b, cu(t), cu(rd), cu(nz) if std > 0 else None, std, white, blender, coef)
    names = ("rgb_map", "disp", "acc", "weights", "depth", "cdisp", "mus", "sigmas", "regs")
    for o, r, nme in zip(out, ref, names):
        if nme in ("disp", "cdisp"):
            close(1.0 / o, 1.0 / r, 1e-4, 1e-5)
        else:
            close(o, r, 1e-4, 2e-6)
    sum((o * cu(c)).sum() for o, c in zip(out, cts)).backward()
    gr, gw = b.grad.cpu(), a.grad
    assert torch.isfinite(gr).all()
    sc = gw.abs().amax(dim=(1, 2), keepdim=True).clamp(min=1e-6)
    err = ((gr - gw).abs() / sc).max().item()
    assert err < 5e-4, err
    # only some cotangents present (what a training step sends: rgb, weights, mus, sigmas, regs)
    b2 = cu(raw6).clone().requires_grad_(True)
    o2 = ops.composite_dd(b2, cu(t), cu(rd), cu(nz) if std > 0 else None, std, white, blender, coef)
    pick = (0, 3, 6, 7, 8)
    sum((o2[i] * cu(cts[i])).sum() for i in pick).backward()
    a2 = raw6.clone().requires_grad_(True)
    r2 = _dd_coarse_reference(a2, t, rd, nz, std, white, blender, coef)
    sum((r2[i] * cts[i]).sum() for i in pick).backward()
    sc = a2.grad.abs().amax(dim=(1, 2), keepdim=True).clamp(min=1e-6)
    assert ((b2.grad.cpu() - a2.grad).abs() / sc).max().item() < 5e-4


def test_composite_dd_regs_are_bit_reproducible(ops):
    """The regulariser sums are reduced in a fixed order (block partials, then the last block): same bits every launch."""
    N, S = 5000, 32
    g = torch.Generator().manual_seed(3)
    t = torch.sort(torch.rand(N, S + 1, generator=g) * 4 + 2, dim=-1)[0]
    raw6, rd = cu(torch.randn(N, S, 6, generator=g)), cu(torch.randn(N, 3, generator=g))
    regs = [ops.composite_dd(raw6, cu(t), rd, None, 0.0, False, True, 0.03)[8].clone() for _ in range(5)]
    for r in regs[1:]:
        assert torch.equal(r, regs[0])
    want = torch.stack([(raw6[..., 4].double() ** 2).sum() / N, (raw6[..., 5].double() ** 2).sum() / N])
    close(regs[0][:2].double(), want, 1e-5, 1e-6)
    close(regs[0][2:], regs[0][:2] * 0.03, 1e-6, 1e-7)


@pytest.mark.parametrize("S,n", [(1, 9), (16, 17), (32, 33), (64, 65), (128, 129), (300, 301)])
@pytest.mark.parametrize("det", [True, False])
def test_fused_dd_resampler_equals_reference_sequence(ops, S, n, det):
    """ddnerf_sample_pdf_mu_sigma_fused (unsmoothed sigmas + gaussian_smooth_factor in, tails evaluated per cell) against
    models.py:268-273 + samplers.py:124-215 done step by step; the factor as a python float and as a device scalar."""
    N = 53
    g, bins, w, mus, sig, _, _ = _resample_inputs(N, S, 400 + S, peaked=(S % 2 == 0))
    smooth = 1.37
    ssig = sig * smooth
    slt = orc.normal_cdf((0 - mus) / ssig)
    spin = orc.normal_cdf((1 - mus) / ssig) - slt
    rand = None if det else torch.rand(N, n, generator=g)
    for pad in (True, False):
        ref, _ = orc.sample_pdf_with_mu_sigma(bins, w, mus, ssig, spin, slt, n, pad, 2.0, 6.0, rand)
        unfused = ops.sample_pdf_mu_sigma(cu(bins), cu(w), cu(mus), cu(ssig), cu(spin), cu(slt), n, pad, 2.0, 6.0, cu(rand))
        for sm in (smooth, torch.tensor(smooth, device=DEV)):
            got = ops.sample_pdf_mu_sigma_fused(cu(bins), cu(w), cu(mus), cu(sig), sm, n, pad, 2.0, 6.0, cu(rand))
            assert (got[:, 1:] >= got[:, :-1]).all()
            # the in-kernel tails follow the reference's fp32 formula: same samples as the unfused kernel up to the erf
            # implementation (CPU torch.erf vs CUDA erff, 1 ulp) amplified by Phi^-1 near the clamps
            err = (got - unfused).abs()
            assert (err > 3e-4).float().mean().item() < 2e-3 and err.max().item() < 5e-2
            err = (got.cpu() - ref).abs()
            assert (err > 3e-4 + 1e-4 * ref.abs()).float().mean().item() < 2e-3 and err.max().item() < 5e-2


@pytest.mark.parametrize("S0,S1", [(3, 5), (16, 16), (32, 32), (64, 64), (128, 128), (300, 300)])
def test_dp_loss_with_in_kernel_tails(ops, S0, S1):
    """dp-loss with left_tails_0 = part_inside_0 = None (evaluated per cell in the kernels) == the same call with the tails
    of models.py:254-258 passed in; forward and backward."""
    N = 41
    g, t0, w0, mus, sig, lt, pin = _resample_inputs(N, S0, 70 + S0, peaked=False)
    w0, sig = w0 + 0.05, sig + 0.15
    lt = orc.normal_cdf((0 - mus) / sig)
    pin = orc.normal_cdf((1 - mus) / sig) - lt
    t1 = torch.sort(torch.rand(N, S1 + 1, generator=g) * 3.8 + 2, dim=-1)[0]
    t1[:, 0] = 2.0
    w1 = torch.rand(N, S1, generator=g) + 0.05
    res = []
    for tails in ((cu(lt), cu(pin)), (None, None)):
        a = [cu(x).clone().requires_grad_(True) for x in (w0, mus, sig)]
        loss = ops.dp_loss(cu(t1), cu(t0), cu(w1), a[0], a[1], a[2], tails[0], tails[1], False)
        loss.backward()
        res.append((loss.detach(), [x.grad for x in a]))
    close(res[1][0], res[0][0], 2e-5, 1e-7)
    # the tails differ by an ulp (CUDA erff in the kernel, CPU torch.erf in the tensors passed in); F = (Phi(x) - lt) / pin
    # cancels where a fine edge sits at a cell's start, so single elements of near-zero gradient rows move visibly
    for g1, g0 in zip(res[1][1], res[0][1]):
        sc = g0.abs().amax(dim=1, keepdim=True).clamp(min=1e-7)
        assert ((g1 - g0).abs() / sc).max().item() < 5e-2
        assert torch.nn.functional.cosine_similarity(g1.flatten().double(), g0.flatten().double(), dim=0).item() > 0.99999


@pytest.mark.parametrize("S0,S1", [(16, 16), (32, 32), (128, 128), (300, 300)])
def test_dp_loss_total_equals_reference_sequence(ops, S0, S1):
    """ops.dp_loss_total == models.py:287-289 done with tensor ops on ops.dp_loss (kl * S1 + mus_reg + sig_reg, [1]):
    forward bit-exact, cotangents of pdf_0 / mus_0 / sigmas_0 and of regs equal."""
    N = 37
    g, t0, w0, mus, sig, _, _ = _resample_inputs(N, S0, 90 + S0, peaked=False)
    w0, sig = w0 + 0.05, sig + 0.15
    t1 = torch.sort(torch.rand(N, S1 + 1, generator=g) * 3.8 + 2, dim=-1)[0]
    t1[:, 0] = 2.0
    w1 = torch.rand(N, S1, generator=g) + 0.05
    regs0 = torch.tensor([0.7, 1.3, 0.021, 0.039])
    res = []
    for fused in (False, True):
        a = [cu(x).clone().requires_grad_(True) for x in (w0, mus, sig, regs0)]
        if fused:
            total = ops.dp_loss_total(cu(t1), cu(t0), cu(w1), a[0], a[1], a[2], None, None, False, a[3], S1)
        else:
            kl = ops.dp_loss(cu(t1), cu(t0), cu(w1), a[0], a[1], a[2], None, None, False) * S1
            total = (kl + a[3][2:3] + a[3][3:4])
        assert total.shape == (1,)
        (total.reshape(()) * 0.37).backward()
        res.append((total.detach().cpu(), [x.grad.cpu() for x in a]))
    if S0 <= 256:
        assert torch.equal(res[0][0], res[1][0])
    else:                                   # the generic kernels (S > 256) sum the per-ray values with atomics: order varies
        close(res[1][0], res[0][0], 1e-6, 0.0)
    for g1, g0 in zip(res[1][1][:3], res[0][1][:3]):
        close(g1, g0, 2e-6, 1e-9)
    assert torch.equal(res[1][1][3], torch.tensor([0.0, 0.0, 0.37, 0.37]))
    assert torch.equal(res[0][1][3], res[1][1][3])


# ---------------------------------------------------------------------------------------------
# K5 depth-distribution loss
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", ["16_16", "32_32", "24_40"])
@pytest.mark.parametrize("wname", ["uniform", "peaked", "bumpy"])
@pytest.mark.parametrize("cname", ["blender", "real"])
def test_dp_loss_golden(ops, shape, wname, cname):
    g = load_golden(f"dp_loss_{shape}")
    w0 = cu(g[f"{wname}_w0"]).requires_grad_(True)
    mus = cu(g[f"{wname}_mus"]).requires_grad_(True)
    sig = cu(g[f"{wname}_sigmas"]).requires_grad_(True)
    loss = ops.dp_loss(cu(g[f"{wname}_t1"]), cu(g["t0"]), cu(g[f"{wname}_w1"]), w0, mus, sig, cu(g[f"{wname}_lt"]),
                       cu(g[f"{wname}_pin"]), cname == "blender")
    if wname == "peaked":
        # fine mass where the coarse estimate is ~0: q = E[k+1]-E[k] is either 0 or one ulp of the CDF
        # (1e-12 vs 6e-8 after the epsilon), a rounding coin flip inside log(); loss agrees to ~1e-3,
        # gradients (~p1/q) are ill-conditioned in the reference itself (see make_golden.py)
        close(loss, g[f"{wname}_{cname}_loss"], 5e-3, 1e-6)
        return
    close(loss, g[f"{wname}_{cname}_loss"], 2e-4, 1e-6)
    loss.backward()
    # Knife edge in the reference itself: the last fine edge sits exactly on `far`, where the estimated
    # CDF equals sum(pdf_0) = 1 +- 1 ulp, and dd_utils.py:66 (est_cdf[est_cdf > 1] = 1) blocks its
    # gradient only when rounding lands above 1 -- a per-ray coin flip that differs between CPU torch,
    # CUDA torch and this kernel.  That edge feeds only the LAST coarse cell's mu/sigma (and a small
    # uniform term of w0 through the cumsum), so the last cell is excluded and w0 gets a looser bound.
    for got, key, tol in ((w0.grad, "g_w0", 3e-2), (mus.grad[:, :-1], "g_mus", 2e-3), (sig.grad[:, :-1], "g_sigmas", 2e-3)):
        ref = g[f"{wname}_{cname}_{key}"]
        ref = ref if key == "g_w0" else ref[:, :-1]
        sc = ref.abs().amax(dim=1, keepdim=True).clamp(min=1e-7)
        assert ((got.cpu() - ref).abs() / sc).max().item() < tol, key


@pytest.mark.parametrize("S0,S1", [(2, 2), (3, 5), (7, 9), (16, 16), (33, 31), (64, 64), (100, 90), (128, 128), (128, 300),
                                   (200, 210), (256, 256), (300, 300)])
@pytest.mark.parametrize("blender", [True, False])
def test_dp_loss_shapes_vs_oracle(ops, S0, S1, blender):
    """Every lane-group shape of the fast path, ragged sizes and the generic fallback, forward and backward."""
    N = 41
    g, t0, w0, mus, sig, lt, pin = _resample_inputs(N, S0, 7 + S0 + S1, peaked=False)
    w0 = w0 + 0.05
    # wide in-cell Gaussians: with sigma ~ 1e-3 the estimated mass of most fine cells is 0 or one ulp of the CDF,
    # a rounding coin flip inside log() in the reference itself (see test_dp_loss_golden, 'peaked')
    sig = sig + 0.15
    lt = orc.normal_cdf((0 - mus) / sig)
    pin = orc.normal_cdf((1 - mus) / sig) - lt
    # fine edges stop short of `far`: at t1 == far the estimated CDF is 1 +- 1 ulp and dd_utils.py:66 blocks or passes
    # its gradient by a rounding coin flip in the reference itself (test_dp_loss_golden documents that knife edge)
    t1 = torch.sort(torch.rand(N, S1 + 1, generator=g) * 3.8 + 2, dim=-1)[0]
    t1[:, 0] = 2.0
    w1 = torch.rand(N, S1, generator=g) + 0.05
    if blender:
        w1[3] = 0.0                                   # a ray the blender row filter drops (dd_utils.py:16)
        lt = lt.clone()                               # the reference forgets to filter left_tails:
        lt[3:] = lt[3:].roll(-1, 0)                   # feed it the rows it would actually pair up
    a = [x.clone().requires_grad_(True) for x in (w0, mus, sig)]
    ref = orc.estimate_dp_loss(t1, t0, w1, a[0], a[1], a[2], lt, pin, blender)
    ref.backward()
    # the same restatement in double: the yardstick for how far fp32 arithmetic alone moves these gradients
    d = [x.double().clone().requires_grad_(True) for x in (w0, mus, sig)]
    orc.estimate_dp_loss(t1.double(), t0.double(), w1.double(), d[0], d[1], d[2], lt.double(), pin.double(), blender).backward()
    if blender:
        lt_ours = lt.clone(); lt_ours[3:] = lt[3:].roll(1, 0)
    else:
        lt_ours = lt
    b = [cu(x).clone().requires_grad_(True) for x in (w0, mus, sig)]
    loss = ops.dp_loss(cu(t1), cu(t0), cu(w1), b[0], b[1], b[2], cu(lt_ours), cu(pin), blender)
    close(loss, ref, 3e-4, 1e-6)
    loss.backward()
    for got, want, exact, tol in zip(b, a, d, (3e-2, 3e-3, 3e-3)):
        gr, gw, gd = got.grad.cpu(), want.grad, exact.grad.float()
        sc = gd.abs().amax(dim=1, keepdim=True).clamp(min=1e-7)
        ours, theirs = ((gr - gd).abs() / sc).max().item(), ((gw - gd).abs() / sc).max().item()
        # as accurate as the reference's own fp32 evaluation (which loses up to a few percent of a row's scale at
        # S0 >= 128), or within the per-kernel budget
        assert ours < max(tol, 3.0 * theirs), (ours, theirs)


# ---------------------------------------------------------------------------------------------
# end to end through the reference-facing model API
# ---------------------------------------------------------------------------------------------
def _build_model(spec, cfg):
    from ddnerf_b200.config import make_cfg
    from ddnerf_b200.models import models as M
    c = make_cfg(model=spec["model"], dataset_type="blender" if spec["blender"] else "REAL360", near=cfg.near,
                 far=cfg.far, num_coarse=spec["nc"], num_fine=spec["nf"], pdf_padding=spec["pad"],
                 gaussian_smooth_factor=spec["smooth"], dist_reg_coeficient=cfg.dist_reg_coeficient)
    model = getattr(M, spec["model"])(c)
    is_dd = spec["model"] == "DDNerfModel"
    model.coarse.load_state_dict(orc.init_mlp_params(is_dd, seed=31))
    if is_dd:
        model.fine.load_state_dict(orc.init_mlp_params(False, seed=32))
    model.to(DEV)
    return model, c


@pytest.mark.parametrize("tag", list(E2E))
def test_end_to_end_golden(ops, tag):
    spec, g, cfg, pc, pf, rnd, rays = e2e_setup(tag)
    model, c = _build_model(spec, cfg)
    model.randoms = {k: cu(v) for k, v in rnd.items()}
    N = g["ro"].shape[0]
    if spec["train"]:
        model.train()
        out = model.run_iter(cu(g["ro"]), cu(g["rd"]), cu(g["rad"]), mode="train", rgb_target=cu(g["target"]))
        target = cu(g["target"])
        loss = 0
        for j in range(2):
            loss = loss + cfg.loss_coeficients[j] * torch.nn.functional.mse_loss(out[j]["rgb"], target)
        if spec["model"] == "DDNerfModel":
            loss = loss + cfg.dp_coeficient * out[1]["dp_loss"].mean()
        loss.backward()
        close(loss, g["loss"], 1e-4, 1e-5)
        for prefix, net in (("gc_", model.coarse), ("gf_", model.fine if spec["model"] == "DDNerfModel" else None)):
            if net is None:
                continue
            for k, p in net.named_parameters():
                ref = g[prefix + k]
                got = p.grad.cpu()
                got = got if got.numel() <= 4096 else got.flatten()[::97]
                sc = ref.abs().max().clamp(min=1e-8)
                # coarse DDNeRF grads inherit the dp-loss knife edge described in test_dp_loss_golden
                tol = 3e-2 if (prefix == "gc_" and spec["model"] == "DDNerfModel") else 5e-3
                assert ((got - ref).abs().max() / sc).item() < tol, (k, ((got - ref).abs().max() / sc).item())
    else:
        model.eval()
        with torch.no_grad():
            out = model.run_iter(cu(g["ro"]).view(N // 8, 8, 3), cu(g["rd"]).view(N // 8, 8, 3),
                                 cu(g["rad"]).view(N // 8, 8, 1), mode="validation")
    for j in range(2):
        for k in ("rgb", "acc", "weights", "depth", "dp_loss", "mus_loss", "sig_loss"):
            key = f"out{j}_{k}"
            if key in g:
                close(out[j][k].reshape(g[key].shape), g[key], 1e-3, 1e-3)      # BASELINE.json fp32 budget
                close(out[j][k].reshape(g[key].shape), g[key], 2e-4, 2e-4)      # what we actually hold
        for k in ("disp", "corrected_disp_map"):
            key = f"out{j}_{k}"
            if key in g:
                close(1.0 / out[j][k].reshape(g[key].shape), 1.0 / g[key], 2e-4, 2e-4)


# ---------------------------------------------------------------------------------------------
# full-size, size-independent properties (BASELINE.json cfg2: 4096 rays x 128+128 samples)
# ---------------------------------------------------------------------------------------------
def test_full_size_properties(ops):
    from ddnerf_b200.rays import synth_rays
    N, S = 4096, 128
    ro, rd, rad, near, far = synth_rays("blender", N, seed=1)
    rays = cu(orc.pack_rays(ro, rd, rad, near, far))
    g = torch.Generator(device=DEV).manual_seed(0)
    t0 = ops.sample_first_cycle(rays[:, 7:8], rays[:, 8:9], S, False, torch.rand(N, S + 1, device=DEV, generator=g))
    assert (t0[:, 1:] >= t0[:, :-1]).all() and (t0[:, 0] == near).all() and (t0[:, -1] == far).all()
    raw = torch.randn(N, S, 4, device=DEV, generator=g) * 3
    rgb, disp, acc, w, depth, _, _ = ops.composite(raw, t0, rays[:, 3:6], None, 0.0, None, False, True, False)
    assert (w >= 0).all() and (acc <= 1 + 1e-5).all()
    assert (rgb >= -0.0011).all() and (rgb <= 1.0011).all()
    assert ((depth >= near - 1e-4) & (depth <= far + 1e-4)).all()
    t1 = ops.sample_pdf(t0, w, S + 1, True, torch.rand(N, S + 1, device=DEV, generator=g))
    assert (t1[:, 1:] >= t1[:, :-1]).all() and (t1 >= near).all() and (t1 <= far).all()
    mus = torch.rand(N, S, device=DEV, generator=g)
    sig = torch.rand(N, S, device=DEV, generator=g) * 0.5 + 1e-3
    lt = 0.5 * (1 + torch.erf((0 - mus) / sig / 2 ** 0.5))
    pin = 0.5 * (1 + torch.erf((1 - mus) / sig / 2 ** 0.5)) - lt
    t2 = ops.sample_pdf_mu_sigma(t0, w, mus, sig, pin, lt, S + 1, False, near, far,
                                 torch.rand(N, S + 1, device=DEV, generator=g))
    assert (t2[:, 1:] >= t2[:, :-1]).all() and (t2[:, 0] == near).all() and (t2[:, -1] == far).all()
    # linearity of the compositor in its cotangent: bwd(2g) == 2 bwd(g)
    rawg = raw.clone().requires_grad_(True)
    o = ops.composite(rawg, t0, rays[:, 3:6], None, 0.0, None, False, True, False)
    ct = torch.randn(N, 3, device=DEV, generator=g)
    (g1,) = torch.autograd.grad((o[0] * ct).sum(), rawg, retain_graph=True)
    (g2,) = torch.autograd.grad((o[0] * ct * 2).sum(), rawg)
    close(g2, 2 * g1, 1e-6, 1e-9)


@pytest.mark.parametrize("N", [1, 37, 4096])
def test_pack_rays_vs_reference_expression(N):
    """a2, get_rays_batches (models.py:144-158): the one-launch ray packing against the reference's
    norm / div / ones_like / cat on the CPU: origins, directions, radius, near, far bit-exact, the
    normalised view directions within one ulp (summation order of the three squares)."""
    from ddnerf_b200 import ops
    g = torch.Generator().manual_seed(N)
    ro, rd = torch.randn(N, 3, generator=g), torch.randn(N, 3, generator=g) * 3.0
    rad = torch.rand(N, 1, generator=g) * 1e-3
    near, far = 2.0, 6.0
    viewdirs = rd / rd.norm(p=2, dim=-1).unsqueeze(-1)
    ref = torch.cat((ro, rd, rad, near * torch.ones_like(rd[..., :1]), far * torch.ones_like(rd[..., :1]), viewdirs), dim=-1)
    out = ops.pack_rays(ro.cuda(), rd.cuda(), rad.cuda(), near, far).cpu()
    assert out.shape == (N, 12)
    assert torch.equal(out[:, :9], ref[:, :9])
    assert (out[:, 9:] - ref[:, 9:]).abs().max().item() <= 1.2e-7


# ---------------------------------------------------------------------------------------------
# edge cases of the boundary: empty ray batches, the largest sizes the entry points accept
# ---------------------------------------------------------------------------------------------
def test_zero_rays_every_operator(ops):
    """N = 0 rays (the last, empty chunk of a sharded frame; a rank with no rows): every operator returns correctly shaped
    empty tensors without launching a kernel (a grid of 0 blocks is a launch error), the scalar losses are 0, and autograd
    through them works."""
    S, n = 8, 9
    z = lambda *shape: torch.zeros(*shape, device=DEV)
    assert ops.sample_first_cycle(z(0, 1), z(0, 1), S, False, z(0, S + 1)).shape == (0, S + 1)
    assert ops.sample_pdf(z(0, S + 1), z(0, S), n, True, z(0, n)).shape == (0, n)
    assert ops.sample_pdf_mu_sigma(z(0, S + 1), z(0, S), z(0, S), z(0, S), z(0, S), z(0, S), n, True, 2.0, 6.0, z(0, n)).shape == (0, n)
    assert ops.sample_pdf_mu_sigma_fused(z(0, S + 1), z(0, S), z(0, S), z(0, S), 1.4, n, True, 2.0, 6.0, z(0, n)).shape == (0, n)
    assert ops.find_interval(z(0, S + 1), z(0, n)).shape == (0, n)
    assert ops.pack_rays(z(0, 3), z(0, 3), z(0, 1), 2.0, 6.0).shape == (0, 12)
    assert ops.encode(z(0, 12), z(0, S + 1)).shape == (0, 123)
    raw = z(0, S, 4).requires_grad_(True)
    out = ops.composite(raw, z(0, S + 1), z(0, 3), z(0, S), 1.0, None, False, True, True)
    assert out[0].shape == (0, 3) and out[3].shape == (0, S) and out[6].shape == (0, S, 3)
    (out[0].sum() + out[3].sum()).backward()
    assert raw.grad.shape == (0, S, 4)
    raw6 = z(0, S, 6).requires_grad_(True)
    dd = ops.composite_dd(raw6, z(0, S + 1), z(0, 3), z(0, S), 1.0, False, True, 0.03)
    assert dd[3].shape == (0, S) and dd[6].shape == (0, S) and dd[8].shape == (4,)
    assert dd[8].abs().max().item() == 0.0
    w0, mu, sg = (z(0, S).requires_grad_(True) for _ in range(3))
    kl = ops.dp_loss(z(0, n), z(0, S + 1), z(0, S), w0, mu, sg, None, None, True)
    assert kl.item() == 0.0
    regs = torch.tensor([0.5, 0.25, 0.0625, 0.125], device=DEV, requires_grad=True)
    tot = ops.dp_loss_total(z(0, n), z(0, S + 1), z(0, S), w0, mu, sg, None, None, True, regs, S)
    assert tot.shape == (1,) and tot.item() == 0.0625 + 0.125
    tot.sum().backward()
    assert w0.grad.shape == (0, S) and torch.equal(regs.grad.cpu(), torch.tensor([0.0, 0.0, 1.0, 1.0]))
    torch.cuda.synchronize()


def test_maximum_sizes_vs_oracle(ops):
    """The largest shapes the entry points accept (include/ddnerf_b200.h): resamplers at S = 2048 cells / 2049 samples,
    compositor at S = 512, dp-loss at 1024 x 1024 -- the generic kernels behind the lane-group fast paths."""
    N = 3
    g, bins, w, mus, sig, lt, pin = _resample_inputs(N, 2048, 5, peaked=True)
    sig = sig + 0.05
    lt = orc.normal_cdf((0 - mus) / sig)
    pin = orc.normal_cdf((1 - mus) / sig) - lt
    for det in (True, False):
        rand = None if det else torch.rand(N, 2049, generator=g)
        s_ref, _ = orc.sample_pdf(bins, w, 2049, True, rand)
        close(ops.sample_pdf(cu(bins), cu(w), 2049, True, cu(rand)), s_ref, 1e-5, 2e-5)
        d_ref, _ = orc.sample_pdf_with_mu_sigma(bins, w, mus, sig, pin, lt, 2049, True, 2.0, 6.0, rand)
        d = ops.sample_pdf_mu_sigma(cu(bins), cu(w), cu(mus), cu(sig), cu(pin), cu(lt), 2049, True, 2.0, 6.0, cu(rand))
        assert (d[:, 1:] >= d[:, :-1]).all()
        err = (d.cpu() - d_ref).abs()
        assert (err > 3e-4 + 1e-4 * d_ref.abs()).float().mean().item() < 2e-3 and err.max().item() < 5e-2
    with pytest.raises(RuntimeError):
        ops.sample_pdf(cu(torch.zeros(1, 2050)), cu(torch.ones(1, 2049)), 9, True, None)          # S = 2049 is refused
    with pytest.raises(RuntimeError):
        ops.composite(cu(torch.zeros(1, 513, 4)), cu(torch.zeros(1, 514)), cu(torch.ones(1, 3)), None, 0.0, None, False, True, False)
    # dp-loss, 1024 coarse x 1024 fine cells
    g, t0, w0, mus, sig, _, _ = _resample_inputs(N, 1024, 6, peaked=False)
    w0, sig = w0 + 0.05, sig + 0.15
    lt = orc.normal_cdf((0 - mus) / sig)
    pin = orc.normal_cdf((1 - mus) / sig) - lt
    t1 = torch.sort(torch.rand(N, 1025, generator=g) * 3.8 + 2, dim=-1)[0]
    t1[:, 0] = 2.0
    w1 = torch.rand(N, 1024, generator=g) + 0.05
    ref = orc.estimate_dp_loss(t1, t0, w1, w0, mus, sig, lt, pin, False)
    close(ops.dp_loss(cu(t1), cu(t0), cu(w1), cu(w0), cu(mus), cu(sig), cu(lt), cu(pin), False), ref, 5e-4, 1e-6)


@pytest.mark.parametrize("pname", ["config_blender", "config_blender_mipnerf"])
def test_run_iter_ragged_chunks_equal_one_chunk(ops, pname):
    """models.py:40-73: run_iter walks the rays in chunks of cfg.nerf[mode].chunksize and concatenates the per-chunk dicts.
    A 5 x 7 validation frame in chunks of 16, 16 and 3 rays equals the same frame in one chunk (rays are independent), the
    image shapes are restored, and in train mode the per-chunk scalars come back as [n_chunks] vectors (the caller takes
    the mean, train_model.py:165)."""
    from ddnerf_b200.config import preset
    from ddnerf_b200.models import models as M
    from ddnerf_b200.rays import synth_rays
    H, W, S = 5, 7, 16
    N = H * W
    ro, rd, rad, _, _ = synth_rays("blender", N, seed=12)
    g = torch.Generator().manual_seed(4)
    rnd = dict(t_rand=torch.rand(N, S + 1, generator=g), noise0=torch.randn(N, S, generator=g),
               u_rand=torch.rand(N, S + 1, generator=g), noise1=torch.randn(N, S, generator=g))
    outs = []
    for chunk in (16, 4096):
        cfg, _ = preset(pname, num_coarse=S, num_fine=S, chunksize=chunk)
        is_dd = cfg.nerf.type == "DDNerfModel"
        model = getattr(M, cfg.nerf.type)(cfg)
        model.coarse.load_state_dict(orc.init_mlp_params(is_dd, seed=41))
        if is_dd:
            model.fine.load_state_dict(orc.init_mlp_params(False, seed=42))
        model.to(torch.device(DEV))
        model.randoms = {k: cu(v) for k, v in rnd.items()}
        model.eval()
        with torch.no_grad():
            val = model.run_iter(cu(ro).view(H, W, 3), cu(rd).view(H, W, 3), cu(rad).view(H, W, 1), mode="validation")
        model.train()
        trn = model.run_iter(cu(ro), cu(rd), cu(rad), mode="train", rgb_target=cu(torch.rand(N, 3, generator=g)))
        outs.append((val, trn, is_dd))
    (va, ta, is_dd), (vb, tb, _) = outs
    for j in range(2):
        assert va[j]["rgb"].shape == (H, W, 3) and va[j]["depth"].shape == (H, W) and va[j]["weights"].shape == (N, S)
        for k in ("rgb", "disp", "acc", "depth", "weights"):
            close(va[j][k], vb[j][k], 1e-6, 1e-6)
            close(ta[j][k], tb[j][k], 1e-6, 1e-6)
    if is_dd:
        assert ta[1]["dp_loss"].shape == (3,) and tb[1]["dp_loss"].shape == (1,)
        assert ta[0]["mus_reg"].shape == (3,) and ta[0]["sig_loss"].shape == (3,)
        assert torch.isfinite(ta[1]["dp_loss"]).all()

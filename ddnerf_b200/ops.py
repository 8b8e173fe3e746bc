"""Tensor-level operators of the hot path: torch.autograd.Functions over the C ABI.

PyTorch is used for device memory, streams and autograd bookkeeping only; every operator body is
one call into libddnerf_b200.so.  Non-CUDA tensors raise (there is no CPU fallback).
"""
import ctypes

import torch

from . import _lib


# bench.py sets this to a list to collect (start, end) CUDA-event pairs around every MLP kernel call
MLP_TIMING = None


class _mlp_timer:
    def __init__(self, tag="mlp"):
        self.tag = tag

    def __enter__(self):
        if MLP_TIMING is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if MLP_TIMING is not None and hasattr(self, "e0"):
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            MLP_TIMING.append((self.e0, e1, self.tag))


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _req(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"ddnerf_b200: `{name}` must be a CUDA tensor (the path has no CPU fallback)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _opt(t, name):
    return None if t is None else _req(t, name)


# ---------------------------------------------------------------------------------------------
# K3 samplers (no gradient: the reference wraps their results in a fresh leaf, samplers.py:121,215)
# ---------------------------------------------------------------------------------------------
def sample_first_cycle(near, far, num_coarse, lindisp=False, t_rand=None):
    """near/far: [N,1] (any stride).  Returns t_vals [N, num_coarse+1]."""
    lib = _lib.load()
    if not near.is_cuda:
        raise RuntimeError("ddnerf_b200: `near` must be a CUDA tensor")
    N = near.shape[0]
    if near.dtype != torch.float32 or far.dtype != torch.float32:
        near, far = near.float(), far.float()
    if near.stride(0) != far.stride(0):
        near, far = near.contiguous(), far.contiguous()
    t_rand = _opt(t_rand, "t_rand")
    out = torch.empty(N, num_coarse + 1, device=near.device, dtype=torch.float32)
    _lib.check(lib.ddnerf_sample_first_cycle(_p(near), _p(far), near.stride(0) if N > 1 else 1, _p(t_rand), _p(out), N,
                                             num_coarse, int(bool(lindisp)), _stream()), "sample_first_cycle")
    return out


def sample_pdf(bins, weights, num_samples, pdf_padding, rand=None, return_idx=False):
    lib = _lib.load()
    bins, weights, rand = _req(bins, "bins"), _req(weights.detach(), "weights"), _opt(rand, "rand")
    N, S = weights.shape
    out = torch.empty(N, num_samples, device=bins.device, dtype=torch.float32)
    idx = torch.empty(N, num_samples, device=bins.device, dtype=torch.int32) if return_idx else None
    _lib.check(lib.ddnerf_sample_pdf(_p(bins), _p(weights), _p(rand), _p(out), _p(idx), N, S, num_samples,
                                     int(bool(pdf_padding)), _stream()), "sample_pdf")
    return (out, idx) if return_idx else out


def sample_pdf_mu_sigma(bins, weights, mus, sigmas, part_inside, left_tail, num_samples, pdf_padding, near_cfg, far_cfg,
                        rand=None, return_idx=False):
    lib = _lib.load()
    bins, weights = _req(bins, "bins"), _req(weights.detach(), "weights")
    mus, sigmas = _req(mus.detach(), "mus"), _req(sigmas.detach(), "sigmas")
    part_inside, left_tail = _req(part_inside.detach(), "part_inside"), _req(left_tail.detach(), "left_tail")
    rand = _opt(rand, "rand")
    N, S = weights.shape
    out = torch.empty(N, num_samples, device=bins.device, dtype=torch.float32)
    idx = torch.empty(N, num_samples, device=bins.device, dtype=torch.int32) if return_idx else None
    _lib.check(lib.ddnerf_sample_pdf_mu_sigma(_p(bins), _p(weights), _p(mus), _p(sigmas), _p(part_inside), _p(left_tail),
                                              _p(rand), _p(out), _p(idx), N, S, num_samples, int(bool(pdf_padding)),
                                              float(near_cfg), float(far_cfg), _stream()), "sample_pdf_mu_sigma")
    return (out, idx) if return_idx else out


def sample_pdf_mu_sigma_fused(bins, weights, mus, sigmas, smooth, num_samples, pdf_padding, near_cfg, far_cfg, rand=None):
    """sample_pdf_with_mu_sigma fed with the UNSMOOTHED sigmas: the kernel applies ``smooth`` (gaussian_smooth_factor; a
    python float, or a 0-dim CUDA tensor when the step is a replayed CUDA graph) and evaluates the smoothed tails of
    models.py:268-273 per cell."""
    lib = _lib.load()
    bins, weights = _req(bins, "bins"), _req(weights.detach(), "weights")
    mus, sigmas = _req(mus.detach(), "mus"), _req(sigmas.detach(), "sigmas")
    rand = _opt(rand, "rand")
    N, S = weights.shape
    out = torch.empty(N, num_samples, device=bins.device, dtype=torch.float32)
    if isinstance(smooth, torch.Tensor):
        if not (smooth.is_cuda and smooth.dtype == torch.float32):
            raise RuntimeError("ddnerf_b200: a tensor gaussian_smooth_factor must be a fp32 CUDA scalar")
        sm_val, sm_dev = 0.0, _p(smooth)
    else:
        sm_val, sm_dev = float(smooth), None
    _lib.check(lib.ddnerf_sample_pdf_mu_sigma_fused(_p(bins), _p(weights), _p(mus), _p(sigmas), sm_val, sm_dev, _p(rand),
                                                    _p(out), None, N, S, num_samples, int(bool(pdf_padding)), float(near_cfg),
                                                    float(far_cfg), _stream()), "sample_pdf_mu_sigma_fused")
    return out


def find_interval(cdf, u):
    """idx[r,k] = #{cdf[r,:] <= u[r,k]} - 1 (int32), the interval search of samplers.py:106-116."""
    lib = _lib.load()
    cdf, u = _req(cdf, "cdf"), _req(u, "u")
    N, n = u.shape
    idx = torch.empty(N, n, device=u.device, dtype=torch.int32)
    _lib.check(lib.ddnerf_find_interval(_p(cdf), _p(u), _p(idx), N, cdf.shape[1] - 1, n, _stream()), "find_interval")
    return idx


# ---------------------------------------------------------------------------------------------
# K2 encoding
# ---------------------------------------------------------------------------------------------
RAY_SHAPES = {"cone": 0, "cylinder": 1}


def pack_rays(ray_origins, ray_directions, ray_rad, near, far):
    """get_rays_batches (models.py:144-158): [N,12] rays = (o, d, radius, near, far, d / ||d||) in one launch."""
    lib = _lib.load()
    ro = _req(ray_origins, "ray_origins").reshape(-1, 3)
    rd = _req(ray_directions, "ray_directions").reshape(-1, 3)
    rad = _req(ray_rad, "ray_rad").reshape(-1)
    N = ro.shape[0]
    if rd.shape[0] != N or rad.shape[0] != N:
        raise RuntimeError("ddnerf_b200: ray_origins, ray_directions and ray_rad disagree on the number of rays")
    rays = torch.empty(N, 12, device=ro.device, dtype=torch.float32)
    _lib.check(lib.ddnerf_pack_rays(_p(ro), _p(rd), _p(rad), float(near), float(far), N, _p(rays), _stream()), "pack_rays")
    return rays


def encode(rays, t_vals, ray_shape="cone"):
    """rays [N,12], t_vals [N,S+1] -> the [N*S,123] MLP input of models.py:133."""
    lib = _lib.load()
    rays, t_vals = _req(rays, "rays"), _req(t_vals, "t_vals")
    N, S = rays.shape[0], t_vals.shape[1] - 1
    out = torch.empty(N * S, 123, device=rays.device, dtype=torch.float32)
    dir_ptr = ctypes.c_void_p(out.data_ptr() + 96 * 4)
    _lib.check(lib.ddnerf_encode(_p(rays), _p(t_vals), _p(out), 123, dir_ptr, 123, N, S, RAY_SHAPES[ray_shape], _stream()),
               "encode")
    return out


# ---------------------------------------------------------------------------------------------
# K1 MLP
# ---------------------------------------------------------------------------------------------
PARAM_ORDER = [f"layers_xyz.{i}" for i in range(8)] + ["fc_feat", "fc_alpha", "layers_dir.0", "fc_rgb", "fc_mu_sigma"]


def _ptr_table(weights, biases):
    t = _lib.MlpPtrs()
    for i in range(_lib.NPARAMS):
        t.w[i] = weights[i].data_ptr() if i < len(weights) and weights[i] is not None else None
        t.b[i] = biases[i].data_ptr() if i < len(biases) and biases[i] is not None else None
    return t


class _MlpF32(torch.autograd.Function):
    """fp32 MLP over encoded rays (rays/t_vals) or an explicit feature matrix x."""

    @staticmethod
    def forward(ctx, x, rays, t_vals, ray_shape, out_channels, *params):
        lib = _lib.load()
        nw = len(params) // 2
        ws_, bs_ = [_req(p, "weight") for p in params[:nw]], [_req(p, "bias") for p in params[nw:]]
        if x is not None:
            x = _req(x, "x")
            rows, dev = x.shape[0], x.device
        else:
            rays, t_vals = _req(rays, "rays"), _req(t_vals, "t_vals")
            N, S = rays.shape[0], t_vals.shape[1] - 1
            rows, dev = N * S, rays.device
        need_bwd = any(ctx.needs_input_grad)
        ws_bytes = lib.ddnerf_mlp_f32_workspace_bytes(rows)
        work = torch.empty(max(ws_bytes, 4) // 4, device=dev, dtype=torch.float32)
        out = torch.empty(rows, out_channels, device=dev, dtype=torch.float32)
        table = _ptr_table(ws_, bs_)
        with _mlp_timer():
            if x is not None:
                _lib.check(lib.ddnerf_mlp_f32_forward_x(ctypes.byref(table), _p(x), rows, out_channels, _p(out), _p(work),
                                                        _stream()), "mlp_f32_forward_x")
            else:
                _lib.check(lib.ddnerf_mlp_f32_forward(ctypes.byref(table), _p(rays), _p(t_vals), N, S,
                                                      RAY_SHAPES[ray_shape], out_channels, _p(out), _p(work), _stream()),
                           "mlp_f32_forward")
        if need_bwd:
            ctx.work = work
            ctx.save_for_backward(*ws_, *bs_)
            ctx.nw, ctx.rows, ctx.out_channels = nw, rows, out_channels
            ctx.want_dx = x is not None and ctx.needs_input_grad[0]
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        saved = ctx.saved_tensors
        ws_, bs_ = saved[:ctx.nw], saved[ctx.nw:]
        grad_out = _req(grad_out, "grad_out")
        total = sum(w.numel() for w in ws_) + sum(b.numel() for b in bs_)
        flat = torch.zeros(total, device=grad_out.device, dtype=torch.float32)
        gws, gbs, off = [], [], 0
        for w in ws_:
            gws.append(flat[off:off + w.numel()].view(w.shape)); off += w.numel()
        for b in bs_:
            gbs.append(flat[off:off + b.numel()].view(b.shape)); off += b.numel()
        dx = torch.empty(ctx.rows, 123, device=grad_out.device, dtype=torch.float32) if ctx.want_dx else None
        pt, gt = _ptr_table(ws_, bs_), _ptr_table(gws, gbs)
        with _mlp_timer():
            _lib.check(lib.ddnerf_mlp_f32_backward(ctypes.byref(pt), ctypes.byref(gt), _p(grad_out), ctx.rows,
                                                   ctx.out_channels, _p(dx), _p(ctx.work), _stream()), "mlp_f32_backward")
        ctx.work = None
        return (dx, None, None, None, None, *gws, *gbs)


def mlp_f32(params, out_channels, x=None, rays=None, t_vals=None, ray_shape="cone"):
    """params: list of 12 or 13 (weight, bias) pairs in PARAM_ORDER.  Returns [rows, out_channels]."""
    ws_ = [w for w, _ in params]
    bs_ = [b for _, b in params]
    return _MlpF32.apply(x, rays, t_vals, ray_shape, out_channels, *ws_, *bs_)


# ---------------------------------------------------------------------------------------------
# K4 compositing
# ---------------------------------------------------------------------------------------------
class _Composite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, t_vals, rd, noise, noise_std, mus, white_background, blender, want_rgb):
        lib = _lib.load()
        if not raw.is_cuda:
            raise RuntimeError("ddnerf_b200: `radiance_field` must be a CUDA tensor (the path has no CPU fallback)")
        raw = raw.float() if raw.dtype != torch.float32 else raw
        # accept a channel-sliced view [N,S,4] of a wider [N,S,C] tensor without copying
        N, S = raw.shape[0], raw.shape[1]
        if not (raw.stride(2) == 1 and raw.stride(1) >= 4 and raw.stride(0) == S * raw.stride(1)):
            raw = raw.contiguous()
        raw_stride = raw.stride(1) if S > 0 else 4
        t_vals = _req(t_vals, "depth_values")
        if not (rd.is_cuda and rd.dtype == torch.float32 and rd.dim() == 2 and rd.stride(1) == 1):
            rd = _req(rd, "ray_directions")
        noise, mus_c = _opt(noise, "noise"), _opt(mus, "mus")
        dev = raw.device
        rgb_map = torch.empty(N, 3, device=dev)
        disp, acc, depth = torch.empty(N, device=dev), torch.empty(N, device=dev), torch.empty(N, device=dev)
        weights = torch.empty(N, S, device=dev)
        cdisp = torch.empty(N, device=dev) if mus is not None else None
        rgb = torch.empty(N, S, 3, device=dev) if want_rgb else None
        _lib.check(lib.ddnerf_composite_forward(_p(raw), raw_stride, _p(t_vals), _p(rd), rd.stride(0) if N > 1 else 3,
                                                _p(noise), float(noise_std), _p(mus_c), int(bool(white_background)),
                                                int(bool(blender)), _p(rgb_map), _p(disp), _p(acc), _p(weights), _p(depth),
                                                _p(cdisp), _p(rgb), N, S, _stream()), "composite_forward")
        ctx.save_for_backward(raw, t_vals, rd, noise, mus_c)
        ctx.cfg = (raw_stride, float(noise_std), bool(white_background), bool(blender), N, S)
        # outputs nobody differentiates (disp, acc, depth, ... in a training step) arrive as None in backward instead of
        # freshly zero-filled tensors: no fill kernels, and the backward kernel skips their terms
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(*([rgb] if rgb is not None else []))
        return rgb_map, disp, acc, weights, depth, cdisp, rgb

    @staticmethod
    def backward(ctx, g_rgb_map, g_disp, g_acc, g_weights, g_depth, g_cdisp, _g_rgb):
        lib = _lib.load()
        raw, t_vals, rd, noise, mus = ctx.saved_tensors
        raw_stride, noise_std, white, blender, N, S = ctx.cfg
        gs = [None if g is None else g.contiguous().float() for g in (g_rgb_map, g_disp, g_acc, g_weights, g_depth, g_cdisp)]
        if all(g is None for g in gs):
            return (None,) * 9
        g_raw = torch.empty(N, S, 4, device=raw.device, dtype=torch.float32)
        want_mus = mus is not None and ctx.needs_input_grad[5]
        g_mus = torch.empty(N, S, device=raw.device, dtype=torch.float32) if want_mus else None
        _lib.check(lib.ddnerf_composite_backward(_p(raw), raw_stride, _p(t_vals), _p(rd), rd.stride(0) if N > 1 else 3,
                                                 _p(noise), noise_std, _p(mus), int(white), int(blender), *[_p(g) for g in gs],
                                                 _p(g_raw), _p(g_mus), N, S, _stream()), "composite_backward")
        return g_raw, None, None, None, None, g_mus, None, None, None


def composite(raw, t_vals, rd, noise=None, noise_std=0.0, mus=None, white_background=False, blender=False, want_rgb=False):
    """Returns (rgb_map, disp, acc, weights, depth, corrected_disp | None, rgb | None).  Gradients flow
    to `raw` (4 channels) and `mus`; not to depth_values / ray_directions (the reference's callers
    never consume those)."""
    return _Composite.apply(raw, t_vals, rd, noise, noise_std, mus, white_background, blender, want_rgb)


class _CompositeDD(torch.autograd.Function):
    """The coarse pass of DDNerfModel.predict after the network (models.py:242-273) as one kernel each way."""

    @staticmethod
    def forward(ctx, raw6, t_vals, rd, noise, noise_std, white_background, blender, dist_reg_coef):
        lib = _lib.load()
        raw6 = _req(raw6, "radiance_field")
        if raw6.dim() != 3 or raw6.shape[2] != 6:
            raise RuntimeError("ddnerf_b200: composite_dd needs the [N,S,6] output of the DDNeRF coarse network")
        N, S = raw6.shape[0], raw6.shape[1]
        t_vals = _req(t_vals, "depth_values")
        if not (rd.is_cuda and rd.dtype == torch.float32 and rd.dim() == 2 and rd.stride(1) == 1):
            rd = _req(rd, "ray_directions")
        noise = _opt(noise, "noise")
        dev = raw6.device
        rgb_map = torch.empty(N, 3, device=dev)
        disp, acc, depth, cdisp = (torch.empty(N, device=dev) for _ in range(4))
        weights, mus, sigmas = (torch.empty(N, S, device=dev) for _ in range(3))
        regs = torch.empty(4, device=dev)
        scratch = torch.empty(lib.ddnerf_composite_dd_scratch_floats(N), device=dev)
        _lib.check(lib.ddnerf_composite_dd_forward(_p(raw6), _p(t_vals), _p(rd), rd.stride(0) if N > 1 else 3, _p(noise),
                                                   float(noise_std), int(bool(white_background)), int(bool(blender)),
                                                   float(dist_reg_coef), _p(rgb_map), _p(disp), _p(acc), _p(weights), _p(depth),
                                                   _p(cdisp), _p(mus), _p(sigmas), _p(regs), _p(scratch), N, S, _stream()),
                   "composite_dd_forward")
        ctx.save_for_backward(raw6, t_vals, rd, noise)
        ctx.cfg = (float(noise_std), bool(white_background), bool(blender), float(dist_reg_coef), N, S)
        ctx.set_materialize_grads(False)
        return rgb_map, disp, acc, weights, depth, cdisp, mus, sigmas, regs

    @staticmethod
    def backward(ctx, g_rgb_map, g_disp, g_acc, g_weights, g_depth, g_cdisp, g_mus, g_sigmas, g_regs):
        lib = _lib.load()
        raw6, t_vals, rd, noise = ctx.saved_tensors
        noise_std, white, blender, coef, N, S = ctx.cfg
        gs = [None if g is None else g.contiguous().float()
              for g in (g_rgb_map, g_disp, g_acc, g_weights, g_depth, g_cdisp, g_mus, g_sigmas, g_regs)]
        if all(g is None for g in gs):
            return (None,) * 8
        g_raw6 = torch.empty(N, S, 6, device=raw6.device, dtype=torch.float32)
        _lib.check(lib.ddnerf_composite_dd_backward(_p(raw6), _p(t_vals), _p(rd), rd.stride(0) if N > 1 else 3, _p(noise),
                                                    noise_std, int(white), int(blender), coef, *[_p(g) for g in gs],
                                                    _p(g_raw6), N, S, _stream()), "composite_dd_backward")
        return g_raw6, None, None, None, None, None, None, None


def composite_dd(raw6, t_vals, rd, noise=None, noise_std=0.0, white_background=False, blender=False, dist_reg_coef=0.0):
    """Returns (rgb_map, disp, acc, weights, depth [corrected], corrected_disp, mus, sigmas, regs[4]) with
    regs = [mus_loss, sig_loss, mus_reg, sig_reg] (models.py:242-264).  Gradients flow to raw6 from every output."""
    return _CompositeDD.apply(raw6, t_vals, rd, noise, noise_std, white_background, blender, dist_reg_coef)


# ---------------------------------------------------------------------------------------------
# K5 depth-distribution loss
# ---------------------------------------------------------------------------------------------
class _DpLoss(torch.autograd.Function):
    """regs is None: kl_div of dd_utils.py:6-78.  regs [4] (+ scale): the loss term of models.py:287-289,
    kl * scale + regs[2] + regs[3], with the cotangent of regs returned too."""

    @staticmethod
    def forward(ctx, t1, t0, w1, w0, mus0, sig0, lt0, pin0, blender, regs, scale):
        lib = _lib.load()
        t1, t0, w1, w0 = _req(t1, "t_vals_1"), _req(t0, "t_vals_0"), _req(w1, "pdf_1"), _req(w0, "pdf_0")
        mus0, sig0 = _req(mus0, "mus_0"), _req(sig0, "sigmas_0")
        if (lt0 is None) != (pin0 is None):
            raise RuntimeError("ddnerf_b200: dp_loss takes left_tails_0 and part_inside_0 together (both None: computed in the kernel)")
        lt0, pin0 = _opt(lt0, "left_tails_0"), _opt(pin0, "part_inside")
        N, S0, S1 = w0.shape[0], w0.shape[1], w1.shape[1]
        scratch = torch.empty(4 + 2 * N, device=w0.device, dtype=torch.float32)      # header + per-ray KL and relevance
        if regs is None:
            loss = torch.empty((), device=w0.device, dtype=torch.float32)
            _lib.check(lib.ddnerf_dp_loss_forward(_p(t1), _p(t0), _p(w1), _p(w0), _p(mus0), _p(sig0), _p(lt0), _p(pin0),
                                                  int(bool(blender)), _p(loss), _p(scratch), N, S0, S1, _stream()), "dp_loss_forward")
        else:
            regs = _req(regs, "regs")
            if regs.numel() != 4:
                raise RuntimeError("ddnerf_b200: dp_loss_total needs regs = [mus_loss, sig_loss, mus_reg, sig_reg]")
            loss = torch.empty(1, device=w0.device, dtype=torch.float32)
            _lib.check(lib.ddnerf_dp_loss_total_forward(_p(t1), _p(t0), _p(w1), _p(w0), _p(mus0), _p(sig0), _p(lt0), _p(pin0),
                                                        int(bool(blender)), float(scale), _p(regs), _p(loss), _p(scratch), N,
                                                        S0, S1, _stream()), "dp_loss_total_forward")
        ctx.save_for_backward(t1, t0, w1, w0, mus0, sig0, lt0, pin0, scratch)
        ctx.cfg = (bool(blender), N, S0, S1, regs is not None, float(scale))
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        lib = _lib.load()
        t1, t0, w1, w0, mus0, sig0, lt0, pin0, scratch = ctx.saved_tensors
        blender, N, S0, S1, total, scale = ctx.cfg
        g_loss = g_loss.contiguous().float()
        g_w0, g_mu, g_sg = torch.empty_like(w0), torch.empty_like(w0), torch.empty_like(w0)
        if not total:
            _lib.check(lib.ddnerf_dp_loss_backward(_p(t1), _p(t0), _p(w1), _p(w0), _p(mus0), _p(sig0), _p(lt0), _p(pin0),
                                                   int(blender), _p(g_loss), _p(scratch), _p(g_w0), _p(g_mu), _p(g_sg), N, S0, S1,
                                                   _stream()), "dp_loss_backward")
            return None, None, None, g_w0, g_mu, g_sg, None, None, None, None, None
        g_regs = torch.empty(4, device=w0.device, dtype=torch.float32)
        _lib.check(lib.ddnerf_dp_loss_total_backward(_p(t1), _p(t0), _p(w1), _p(w0), _p(mus0), _p(sig0), _p(lt0), _p(pin0),
                                                     int(blender), scale, _p(g_loss), _p(scratch), _p(g_w0), _p(g_mu), _p(g_sg),
                                                     _p(g_regs), N, S0, S1, _stream()), "dp_loss_total_backward")
        return None, None, None, g_w0, g_mu, g_sg, None, None, None, g_regs, None


def dp_loss(t1, t0, pdf_1, pdf_0, mus_0, sigmas_0, left_tails_0, part_inside_0, blender):
    """kl_div(log q, p1, 'mean') of dd_utils.py:6-78.  Gradients flow to pdf_0, mus_0, sigmas_0.  With
    left_tails_0 = part_inside_0 = None the kernels evaluate the two tails themselves (models.py:254-258)."""
    return _DpLoss.apply(t1, t0, pdf_1, pdf_0, mus_0, sigmas_0, left_tails_0, part_inside_0, blender, None, 1.0)


def dp_loss_total(t1, t0, pdf_1, pdf_0, mus_0, sigmas_0, left_tails_0, part_inside_0, blender, regs, scale):
    """models.py:287-289 in one call: ``(estimate_dp_loss(...) * scale + mus_reg + sig_reg).unsqueeze(0)`` with
    ``regs = [mus_loss, sig_loss, mus_reg, sig_reg]`` (the 4-vector composite_dd returns).  Returns a [1] tensor;
    gradients flow to pdf_0, mus_0, sigmas_0 and regs."""
    return _DpLoss.apply(t1, t0, pdf_1, pdf_0, mus_0, sigmas_0, left_tails_0, part_inside_0, blender, regs, scale)


# ---------------------------------------------------------------------------------------------
# training-step tail
# ---------------------------------------------------------------------------------------------
def mse_loss_and_grad(rgb0, rgb1, target, coef0, coef1):
    """Returns (mse[3], g_rgb0, g_rgb1): the photometric terms of train_model.py:159-163 -- mse[0], mse[1] and their
    weighted sum coef0*mse0 + coef1*mse1 in mse[2] -- and the cotangents of that sum w.r.t. rgb0 / rgb1."""
    lib = _lib.load()
    rgb0, target = _req(rgb0.detach(), "rgb0"), _req(target, "target")
    rgb1 = _opt(None if rgb1 is None else rgb1.detach(), "rgb1")
    g0 = torch.empty_like(rgb0)
    g1 = torch.empty_like(rgb1) if rgb1 is not None else None
    mse = torch.empty(3, device=rgb0.device, dtype=torch.float32)
    _lib.check(lib.ddnerf_mse_loss(_p(rgb0), _p(rgb1), _p(target), float(coef0), float(coef1), _p(g0), _p(g1), _p(mse),
                                   rgb0.shape[0], _stream()), "mse_loss")
    return mse, g0, g1


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, step, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    """In-place torch.optim.Adam update of a flat fp32 bucket."""
    lib = _lib.load()
    for t in (param, grad, exp_avg, exp_avg_sq):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError("ddnerf_b200: adam_step needs contiguous fp32 CUDA buffers")
    _lib.check(lib.ddnerf_adam_step(_p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), param.numel(), float(lr), float(beta1),
                                    float(beta2), float(eps), int(step), float(grad_scale), _stream()), "adam_step")


def adam_step_dev(param, grad, exp_avg, exp_avg_sq, hyper):
    """adam_step with {lr, beta1, beta2, eps, 1-beta1^t, sqrt(1-beta2^t), grad_scale, -, 1-beta1, 1-beta2} read from the
    device tensor `hyper` (10 floats, what ddnerf_train_schedule writes): graph-replayable."""
    lib = _lib.load()
    for t in (param, grad, exp_avg, exp_avg_sq, hyper):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError("ddnerf_b200: adam_step_dev needs contiguous fp32 CUDA buffers")
    if hyper.numel() < 10:
        raise RuntimeError("ddnerf_b200: adam_step_dev needs the 10-float hyper record of ddnerf_train_schedule")
    _lib.check(lib.ddnerf_adam_step_dev(_p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), param.numel(), _p(hyper), _stream()),
               "adam_step_dev")

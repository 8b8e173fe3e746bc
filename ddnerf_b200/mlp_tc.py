"""bf16 tensor-core mode of the NeRF MLP: host-side state (packed weight images) and the autograd
Function over the fused tcgen05 kernels of csrc/mlp_tc.cu.  No fallback: the kernels need sm_100."""
import ctypes

import torch

from . import _lib
from .ops import RAY_SHAPES, _mlp_timer, _p, _ptr_table, _req, _stream


class TcState:
    """Packed bf16 weight image + fp32 bias table of one network, refreshed when parameters change."""

    def __init__(self, module):
        self.module = module
        self.key = None
        self.wimg = None
        self.bias = None
        self.dirty = True

    def _params(self):
        return self.module._param_pairs()

    def refresh(self):
        lib = _lib.load()
        pairs = self._params()
        key = tuple((w.data_ptr(), w._version, b.data_ptr(), b._version) for w, b in pairs)
        if not self.dirty and key == self.key and self.wimg is not None:
            return
        dev = pairs[0][0].device
        if self.wimg is None or self.wimg.device != dev:
            self.wimg = torch.empty(lib.ddnerf_mlp_tc_wimg_bytes(), device=dev, dtype=torch.uint8)
            self.bias = torch.empty(lib.ddnerf_mlp_tc_bias_floats(), device=dev, dtype=torch.float32)
        ws_ = [_req(w.detach(), "weight") for w, _ in pairs]
        bs_ = [_req(b.detach(), "bias") for _, b in pairs]
        table = _ptr_table(ws_, bs_)
        _lib.check(lib.ddnerf_mlp_tc_pack(ctypes.byref(table), self.module.out_channels, _p(self.wimg), _p(self.bias), _stream()),
                   "mlp_tc_pack")
        self.key, self.dirty = key, False


def _state(module):
    st = getattr(module, "_tc_state", None)
    if st is None:
        st = TcState(module)
        object.__setattr__(module, "_tc_state", st)
    return st


def encode_img(rays, t_vals, ray_shape="cone"):
    """rays [N,12], t_vals [N,S+1] -> bf16 operand images of all 256-row work items (uint8 tensor), by the STANDALONE
    encoder kernel.  The model path no longer calls it (the chain kernel encodes in its own warps); kept for the layout
    tests and for callers of the image-fed entry point ddnerf_mlp_tc_forward."""
    lib = _lib.load()
    rays, t_vals = _req(rays, "rays"), _req(t_vals.detach(), "t_vals")
    N, S = rays.shape[0], t_vals.shape[1] - 1
    img = torch.empty(max(lib.ddnerf_mlp_tc_enc_bytes(N * S), 16), device=rays.device, dtype=torch.uint8)
    _lib.check(lib.ddnerf_mlp_tc_encode(_p(rays), _p(t_vals), N, S, RAY_SHAPES[ray_shape], _p(img), _stream()), "mlp_tc_encode")
    return img


_ENC_SCRATCH = {}


def _enc_scratch(dev):
    """Per-device double buffer of the in-kernel encoder (2 x 64 KB per SM; lives in L2).  Launches on one stream are
    ordered, so one buffer per (device, stream) serves every inference call."""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    buf = _ENC_SCRATCH.get(key)
    if buf is None:
        buf = torch.empty(_lib.load().ddnerf_mlp_tc_enc_scratch_bytes(), device=dev, dtype=torch.uint8)
        _ENC_SCRATCH[key] = buf
    return buf


def forward_only(module, rays, t_vals, ray_shape="cone"):
    """Inference forward: [N*S, C] fp32 raw network outputs.  ONE launch: the chain kernel's encoder warps compute the
    integrated positional encoding of work item n+1 while the GEMMs of item n run (no encoded image in HBM)."""
    lib = _lib.load()
    st = _state(module)
    st.refresh()
    rays, t_vals = _req(rays, "rays"), _req(t_vals.detach(), "t_vals")
    N, S = rays.shape[0], t_vals.shape[1] - 1
    rows = N * S
    out = torch.empty(rows, module.out_channels, device=rays.device, dtype=torch.float32)
    if rows == 0:
        return out
    scratch = _enc_scratch(rays.device)
    with _mlp_timer():
        _lib.check(lib.ddnerf_mlp_tc_forward_rays(_p(st.wimg), _p(st.bias), _p(rays), _p(t_vals), N, S, RAY_SHAPES[ray_shape],
                                                  module.out_channels, _p(out), None, _p(scratch), None, None, _stream()),
                   "mlp_tc_forward_rays")
    return out


class _MlpTc(torch.autograd.Function):
    """Training path: forward keeps every layer's bf16 activations (tile images) and ReLU masks;
    backward = dX chain kernel + weight-gradient kernel.  No gradient flows to rays / t_vals (the
    reference's sampler output is a detached leaf, models.py:227-237)."""

    @staticmethod
    def forward(ctx, module, rays, t_vals, ray_shape, *params):
        lib = _lib.load()
        st = _state(module)
        st.refresh()
        N, S = rays.shape[0], t_vals.shape[1] - 1
        rows, C, dev = N * S, module.out_channels, rays.device
        # the operand image is written by the forward kernel's encoder warps and kept for the weight-gradient kernel
        img = torch.empty(max(lib.ddnerf_mlp_tc_enc_bytes(rows), 16), device=dev, dtype=torch.uint8)
        out = torch.empty(rows, C, device=dev, dtype=torch.float32)
        act = torch.empty(max(lib.ddnerf_mlp_tc_act_save_bytes(rows), 16), device=dev, dtype=torch.uint8)
        mask = torch.empty(max(lib.ddnerf_mlp_tc_mask_save_bytes(rows), 16), device=dev, dtype=torch.uint8)
        with _mlp_timer("fwd"):
            _lib.check(lib.ddnerf_mlp_tc_forward_rays(_p(st.wimg), _p(st.bias), _p(rays), _p(t_vals), N, S,
                                                      RAY_SHAPES[ray_shape], C, _p(out), _p(img), None, _p(act), _p(mask),
                                                      _stream()), "mlp_tc_forward_rays")
        ctx.st, ctx.img, ctx.act, ctx.mask = st, img, act, mask
        ctx.rows, ctx.C = rows, C
        ctx.sink = getattr(module, "_grad_sink", None)
        ctx.shapes = [tuple(p.shape) for p in params]
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        st, rows, C = ctx.st, ctx.rows, ctx.C
        grad_out = _req(grad_out, "grad_out")
        dev = grad_out.device
        nw = len(ctx.shapes) // 2
        if ctx.sink is not None:
            # the trainer's flat gradient bucket: the dW kernel accumulates straight into it (both passes of a shared
            # network, every chunk), so autograd neither allocates, clones, adds nor concatenates parameter gradients
            gws, gbs = ctx.sink
            views = [None] * len(ctx.shapes)
        else:
            sizes = [int(torch.Size(s).numel()) for s in ctx.shapes]
            flat = torch.zeros(sum(sizes), device=dev, dtype=torch.float32)
            views, off = [], 0
            for s, n in zip(ctx.shapes, sizes):
                views.append(flat[off:off + n].view(s))
                off += n
            gws, gbs = views[:nw], views[nw:]
        dz = torch.empty(ctx.act.numel(), device=dev, dtype=torch.uint8)
        table = _ptr_table(gws, gbs)
        with _mlp_timer("dx"):
            _lib.check(lib.ddnerf_mlp_tc_backward_dx(_p(st.wimg), _p(st.bias), _p(grad_out), rows, C, _p(ctx.mask), _p(dz),
                                                     0, _stream()), "mlp_tc_backward_dx")
        with _mlp_timer("dw"):
            _lib.check(lib.ddnerf_mlp_tc_backward_dw(_p(ctx.act), _p(dz), _p(ctx.img), _p(grad_out), ctypes.byref(table), rows, C,
                                                     0, _stream()), "mlp_tc_backward_dw")
        ctx.img = ctx.act = ctx.mask = None
        return (None, None, None, None, *views)


def mlp_bf16(module, rays, t_vals, ray_shape="cone"):
    needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters())
    if not needs_grad:
        return forward_only(module, rays, t_vals, ray_shape)
    pairs = module._param_pairs()
    rays, t_vals = _req(rays, "rays"), _req(t_vals.detach(), "t_vals")
    return _MlpTc.apply(module, rays, t_vals, ray_shape, *[w for w, _ in pairs], *[b for _, b in pairs])

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
    """rays [N,12], t_vals [N,S+1] -> bf16 operand images of all 256-row work items (uint8 tensor)."""
    lib = _lib.load()
    rays, t_vals = _req(rays, "rays"), _req(t_vals.detach(), "t_vals")
    N, S = rays.shape[0], t_vals.shape[1] - 1
    img = torch.empty(max(lib.ddnerf_mlp_tc_enc_bytes(N * S), 16), device=rays.device, dtype=torch.uint8)
    _lib.check(lib.ddnerf_mlp_tc_encode(_p(rays), _p(t_vals), N, S, RAY_SHAPES[ray_shape], _p(img), _stream()), "mlp_tc_encode")
    return img


def forward_only(module, rays, t_vals, ray_shape="cone"):
    """Inference forward: [N*S, C] fp32 raw network outputs."""
    lib = _lib.load()
    st = _state(module)
    st.refresh()
    N, S = rays.shape[0], t_vals.shape[1] - 1
    rows = N * S
    img = encode_img(rays, t_vals, ray_shape)
    out = torch.empty(rows, module.out_channels, device=rays.device, dtype=torch.float32)
    with _mlp_timer():
        _lib.check(lib.ddnerf_mlp_tc_forward(_p(st.wimg), _p(st.bias), _p(img), rows, module.out_channels, _p(out), None, None,
                                             _stream()), "mlp_tc_forward")
    return out


def mlp_bf16(module, rays, t_vals, ray_shape="cone"):
    needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters())
    if needs_grad:
        raise NotImplementedError("bf16 MLP backward is not wired yet")
    return forward_only(module, rays, t_vals, ray_shape)

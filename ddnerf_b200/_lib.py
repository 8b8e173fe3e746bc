"""ctypes binding of libddnerf_b200.so (the C ABI declared in include/ddnerf_b200.h).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# (DDNERF_B200_LIB: another build of the same ABI, for A/B measurements of kernel variants)
LIB_PATH = os.environ.get("DDNERF_B200_LIB") or os.path.join(_HERE, "libddnerf_b200.so")

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_l = ctypes.c_int64
c_f = ctypes.c_float

NPARAMS = 13


class MlpPtrs(ctypes.Structure):
    """ddnerf_mlp_params / ddnerf_mlp_grads (same layout)."""
    _fields_ = [("w", c_p * NPARAMS), ("b", c_p * NPARAMS)]


# name -> (restype, argtypes); must list every symbol of include/ddnerf_b200.h
SIGNATURES = {
    "ddnerf_version": (c_i, []),
    "ddnerf_last_error": (ctypes.c_char_p, []),
    "ddnerf_device_is_sm100": (c_i, []),
    "ddnerf_launch_count": (c_l, []),
    "ddnerf_sample_first_cycle": (c_i, [c_p, c_p, c_l, c_p, c_p, c_l, c_i, c_i, c_p]),
    "ddnerf_sample_pdf": (c_i, [c_p, c_p, c_p, c_p, c_p, c_l, c_i, c_i, c_i, c_p]),
    "ddnerf_sample_pdf_mu_sigma": (c_i, [c_p] * 9 + [c_l, c_i, c_i, c_i, c_f, c_f, c_p]),
    "ddnerf_ray_bundle": (c_i, [c_i, c_i, ctypes.c_double, c_p, c_i, c_f, c_i, c_i, c_p, c_p, c_p, c_p]),
    "ddnerf_ray_bundle_dev": (c_i, [c_i, c_i, ctypes.c_double, c_p, c_i, c_f, c_i, c_i, c_p, c_p, c_p, c_p]),
    "ddnerf_frame_minmax": (c_i, [c_p, c_l, c_p, c_p, c_p]),
    "ddnerf_frame_pack_u8": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_p]),
    "ddnerf_raystore_pack": (c_i, [c_p, c_p, c_p, c_p, c_l, c_p, c_p]),
    "ddnerf_raystore_gather": (c_i, [c_p, c_l, c_p, c_l, c_l, c_p, c_p, c_p, c_p, c_p, c_p]),
    "ddnerf_find_interval": (c_i, [c_p, c_p, c_p, c_l, c_i, c_i, c_p]),
    "ddnerf_pack_rays": (c_i, [c_p, c_p, c_p, c_f, c_f, c_l, c_p, c_p]),
    "ddnerf_encode": (c_i, [c_p, c_p, c_p, c_l, c_p, c_l, c_l, c_i, c_i, c_p]),
    "ddnerf_mlp_f32_workspace_bytes": (c_l, [c_l]),
    "ddnerf_mlp_f32_forward": (c_i, [ctypes.POINTER(MlpPtrs), c_p, c_p, c_l, c_i, c_i, c_i, c_p, c_p, c_p]),
    "ddnerf_mlp_f32_forward_x": (c_i, [ctypes.POINTER(MlpPtrs), c_p, c_l, c_i, c_p, c_p, c_p]),
    "ddnerf_mlp_f32_backward": (c_i, [ctypes.POINTER(MlpPtrs), ctypes.POINTER(MlpPtrs), c_p, c_l, c_i, c_p, c_p, c_p]),
    "ddnerf_mlp_tc_wimg_bytes": (c_l, []),
    "ddnerf_mlp_tc_bias_floats": (c_l, []),
    "ddnerf_mlp_tc_items": (c_l, [c_l]),
    "ddnerf_mlp_tc_enc_bytes": (c_l, [c_l]),
    "ddnerf_mlp_tc_act_save_bytes": (c_l, [c_l]),
    "ddnerf_mlp_tc_mask_save_bytes": (c_l, [c_l]),
    "ddnerf_mlp_tc_pack": (c_i, [ctypes.POINTER(MlpPtrs), c_i, c_p, c_p, c_p]),
    "ddnerf_mlp_tc_encode": (c_i, [c_p, c_p, c_l, c_i, c_i, c_p, c_p]),
    "ddnerf_mlp_tc_forward": (c_i, [c_p, c_p, c_p, c_l, c_i, c_p, c_p, c_p, c_p]),
    "ddnerf_mlp_tc_enc_scratch_bytes": (c_l, []),
    "ddnerf_mlp_tc_forward_rays": (c_i, [c_p, c_p, c_p, c_p, c_l, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p]),
    "ddnerf_mlp_tc_backward_dx": (c_i, [c_p, c_p, c_p, c_l, c_i, c_p, c_p, c_i, c_p]),
    "ddnerf_mlp_tc_backward_dw": (c_i, [c_p, c_p, c_p, c_p, ctypes.POINTER(MlpPtrs), c_l, c_i, c_i, c_p]),
    "ddnerf_mlp_tc_program_check": (c_i, []),
    "ddnerf_mlp_tc_set_pair_mode": (c_i, [c_i]),
    "ddnerf_mlp_tc_set_profile_buffer": (c_i, [c_p]),
    "ddnerf_mlp_tc_dw_set_profile_buffer": (c_i, [c_p]),
    "ddnerf_mlp_tc_dw_plan": (c_i, [c_l, c_i, ctypes.POINTER(ctypes.c_uint32), c_i]),
    "ddnerf_tc_gemm_selftest": (c_i, [c_p, c_l, c_p, c_l, c_p, c_i, c_i, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32,
                                      ctypes.POINTER(ctypes.c_uint32), c_p]),
    "ddnerf_tc_mma_rate": (c_i, [c_i, c_i, c_i, c_i, c_p, c_p]),
    "ddnerf_composite_forward": (c_i, [c_p, c_i, c_p, c_p, c_l, c_p, c_f, c_p, c_i, c_i] + [c_p] * 7 + [c_l, c_i, c_p]),
    "ddnerf_composite_backward": (c_i, [c_p, c_i, c_p, c_p, c_l, c_p, c_f, c_p, c_i, c_i] + [c_p] * 8 + [c_l, c_i, c_p]),
    "ddnerf_sample_pdf_mu_sigma_fused": (c_i, [c_p] * 4 + [c_f] + [c_p] * 4 + [c_l, c_i, c_i, c_i, c_f, c_f, c_p]),
    "ddnerf_composite_dd_scratch_floats": (c_l, [c_l]),
    "ddnerf_composite_dd_forward": (c_i, [c_p, c_p, c_p, c_l, c_p, c_f, c_i, c_i, c_f] + [c_p] * 10 + [c_l, c_i, c_p]),
    "ddnerf_composite_dd_backward": (c_i, [c_p, c_p, c_p, c_l, c_p, c_f, c_i, c_i, c_f] + [c_p] * 10 + [c_l, c_i, c_p]),
    "ddnerf_dp_loss_forward": (c_i, [c_p] * 8 + [c_i, c_p, c_p, c_l, c_i, c_i, c_p]),
    "ddnerf_dp_loss_backward": (c_i, [c_p] * 8 + [c_i] + [c_p] * 5 + [c_l, c_i, c_i, c_p]),
    "ddnerf_dp_loss_total_forward": (c_i, [c_p] * 8 + [c_i, c_f, c_p, c_p, c_p, c_l, c_i, c_i, c_p]),
    "ddnerf_dp_loss_total_backward": (c_i, [c_p] * 8 + [c_i, c_f] + [c_p] * 6 + [c_l, c_i, c_i, c_p]),
    "ddnerf_mse_loss": (c_i, [c_p, c_p, c_p, c_f, c_f, c_p, c_p, c_p, c_l, c_p]),
    "ddnerf_adam_step": (c_i, [c_p, c_p, c_p, c_p, c_l, c_f, ctypes.c_double, ctypes.c_double, ctypes.c_double, c_i, c_f, c_p]),
    "ddnerf_adam_step_dev": (c_i, [c_p, c_p, c_p, c_p, c_l, c_p, c_p]),
    "ddnerf_train_schedule": (c_i, [c_p, c_p, ctypes.POINTER(ctypes.c_double), c_p]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
            "Build it with `python -m ddnerf_b200.build`.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.ddnerf_version() != 1:
        raise RuntimeError("libddnerf_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().ddnerf_last_error().decode(errors="replace")
        raise RuntimeError(f"ddnerf_b200 {what} failed (code {rc}): {msg}")

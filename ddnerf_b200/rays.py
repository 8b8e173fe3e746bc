"""Host-side synthetic ray generation (dataset-shaped rays for tests and the benchmark).

Restates, on the CPU, the once-per-image ray set-up of the reference so synthetic workloads have
the same distributions as its datasets:

* ``get_ray_bundle``      -- general_utils/nerf_helpers.py:67-125
* ``ndc_mipnerf_rays``    -- data_utils/dataset_helpers.py:3-42
* ``pose_spherical``      -- data_utils/load_blender.py:12-41

Row f1 ("next") of SURVEY.md section 8f: ``ray_bundle_cuda`` is the device kernel (csrc/raygen.cu) the
render loop uses; the tensor-level functions below restate the same arithmetic on the CPU to synthesise
benchmark / test rays and to check the kernel.
"""
import math

import numpy as np
import torch


def pose_spherical(theta_deg, phi_deg, radius):
    """load_blender.py:12-41: camera-to-world for a camera on a sphere looking at the origin."""
    t = np.eye(4, dtype=np.float32)
    t[2, 3] = radius
    phi = phi_deg / 180.0 * np.pi
    rx = np.eye(4, dtype=np.float32)
    rx[1, 1] = rx[2, 2] = np.cos(phi)
    rx[1, 2] = -np.sin(phi)
    rx[2, 1] = np.sin(phi)
    th = theta_deg / 180.0 * np.pi
    ry = np.eye(4, dtype=np.float32)
    ry[0, 0] = ry[2, 2] = np.cos(th)
    ry[0, 2] = -np.sin(th)
    ry[2, 0] = np.sin(th)
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float32)
    return torch.from_numpy((flip @ ry @ rx @ t).astype(np.float32))


def get_ray_bundle(height, width, focal, c2w):
    """nerf_helpers.py:67-125 -> origins [H,W,3], directions [H,W,3] (un-normalised), radii [H,W,1].
    A pose that lives on a CUDA device gets its rays from the device kernel (``ray_bundle_cuda``); a CPU pose
    is the host-side restatement used to synthesise dataset-shaped workloads."""
    if isinstance(c2w, torch.Tensor) and c2w.is_cuda:
        return ray_bundle_cuda(height, width, focal, c2w, c2w.device)
    eps = 1e-5
    jj, ii = torch.meshgrid(torch.arange(height, dtype=c2w.dtype), torch.arange(width, dtype=c2w.dtype),
                            indexing="ij")
    dirs = torch.stack([(ii - width * 0.5) / focal, -(jj - height * 0.5) / focal, -torch.ones_like(ii)], -1)
    rd = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    ro = c2w[:3, -1].expand(rd.shape).clone()
    ro[ro == 0] += eps                                                    # nerf_helpers.py:114-115
    rd[rd == 0] += eps
    dx = torch.sqrt(torch.sum((dirs[:-1] - dirs[1:]) ** 2, -1))
    dx = torch.cat([dx, dx[-2:-1]], 0)
    radii = dx[..., None] * 2 / np.sqrt(12)
    return ro, rd, radii


def ndc_project(H, W, focal, near, rays_o, rays_d):
    """The NDC projection shared by nerf_helpers.py:182-208 (ndc_rays) and dataset_helpers.py:3-42:
    origins moved to the near plane, then x/z, y/z scaled by the focal length, z -> 1 + 2 near / z."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    sx, sy = -1.0 / (W / (2.0 * focal)), -1.0 / (H / (2.0 * focal))
    oz = rays_o[..., 2]
    ox_z, oy_z = rays_o[..., 0] / oz, rays_o[..., 1] / oz
    o = torch.stack([sx * ox_z, sy * oy_z, 1.0 + 2.0 * near / oz], -1)
    d = torch.stack([sx * (rays_d[..., 0] / rays_d[..., 2] - ox_z), sy * (rays_d[..., 1] / rays_d[..., 2] - oy_z),
                     -2.0 * near / oz], -1)
    return o, d


def ndc_mipnerf_rays(H, W, focal, rays_o, rays_d, near=1):
    """dataset_helpers.py:3-42: forward-facing NDC rays + radii from neighbouring origins.  (Tensor-level
    restatement; the fused device path is ``ray_bundle_cuda(..., ndc_near=near)``.)"""
    o, d = ndc_project(H, W, focal, near, rays_o, rays_d)
    dx = torch.sqrt(torch.sum((o[:-1] - o[1:]) ** 2, -1))
    dx = torch.cat([dx, dx[-2:-1]])
    dy = torch.sqrt(torch.sum((o[:, :-1] - o[:, 1:]) ** 2, -1))
    dy = torch.cat([dy, dy[:, -2:-1]], 1)
    radii = (0.5 * (dx + dy)) * 2 / math.sqrt(12)
    return o, d, radii[..., None]


def ray_bundle_cuda(height, width, focal, c2w, device="cuda", rows=None, ndc_near=None):
    """get_ray_bundle (and, with ``ndc_near`` set, ndc_mipnerf_rays on top of it) as ONE kernel on the device:
    the frame's rays from the 12 floats of its pose, no meshgrid and no host->device copy of the rays.
    ``rows=(lo, hi)`` restricts the output to those pixel rows (frame split across ranks).  Returns
    (origins [rows,W,3], directions [rows,W,3], radii [rows,W,1]) as CUDA tensors."""
    import ctypes

    from . import _lib
    lib = _lib.load()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("ddnerf_b200: ray_bundle_cuda needs a CUDA device (the path has no CPU fallback)")
    lo, hi = (0, height) if rows is None else rows
    pose = (ctypes.c_float * 12)(*[float(x) for x in torch.as_tensor(c2w).detach().cpu().reshape(-1)[:12]])
    with torch.cuda.device(dev):
        ro = torch.empty(hi - lo, width, 3, device=dev)
        rd = torch.empty(hi - lo, width, 3, device=dev)
        rad = torch.empty(hi - lo, width, 1, device=dev)
        _lib.check(lib.ddnerf_ray_bundle(height, width, float(focal), pose, int(ndc_near is not None),
                                         float(ndc_near or 0.0), lo, hi, ctypes.c_void_p(ro.data_ptr()),
                                         ctypes.c_void_p(rd.data_ptr()), ctypes.c_void_p(rad.data_ptr()),
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "ray_bundle")
    return ro, rd, rad


# Workload presets: (H, W, focal, c2w, near, far, ndc) -- SURVEY.md section 8d
def frame(kind):
    if kind == "blender":        # config_blender.yml: 800x800 lego-like camera
        H = W = 800
        focal = 0.5 * W / math.tan(0.5 * 0.6911112)
        return H, W, focal, pose_spherical(30.0, -30.0, 4.0), 2.0, 6.0, False
    if kind == "ff":             # config_ff.yml: 1008x756 fern-like NDC camera
        return 756, 1008, 815.0, torch.eye(4), 0.0, 1.0, True
    if kind == "360":            # config_360.yml: normalised bounded-360 scene
        return 756, 1008, 815.0, pose_spherical(40.0, -10.0, 0.89), 1.0 / 5, 14.0 / 5, False
    raise ValueError(kind)


def full_frame_rays(kind):
    """All H*W rays of the preset frame as (ro[H,W,3], rd[H,W,3], rad[H,W,1], near, far)."""
    H, W, focal, c2w, near, far, ndc = frame(kind)
    ro, rd, rad = get_ray_bundle(H, W, focal, c2w)
    if ndc:
        ro, rd, rad = ndc_mipnerf_rays(H, W, focal, ro, rd, near=1)
    return ro.contiguous(), rd.contiguous(), rad.contiguous(), near, far


_FRAME_CACHE = {}


def synth_rays(kind, n, seed=0):
    """``n`` rays drawn without replacement from the preset frame (like dataset.py:56-57).
    Returns (ro[n,3], rd[n,3], rad[n,1], near, far); if n exceeds the frame the draw wraps."""
    if kind not in _FRAME_CACHE:
        _FRAME_CACHE[kind] = full_frame_rays(kind)
    ro, rd, rad, near, far = _FRAME_CACHE[kind]
    total = ro.shape[0] * ro.shape[1]
    g = torch.Generator().manual_seed(1000 + seed)
    if n <= total:
        idx = torch.randperm(total, generator=g)[:n]
    else:
        idx = torch.randint(0, total, (n,), generator=g)
    return (ro.reshape(-1, 3)[idx].contiguous(), rd.reshape(-1, 3)[idx].contiguous(),
            rad.reshape(-1, 1)[idx].contiguous(), near, far)

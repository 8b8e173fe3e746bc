"""Device-resident training ray store: the reference's ``TrainDataset`` (data_utils/dataset.py:8-59) with its rays
in HBM (row f3 of SURVEY.md section 8f).

The reference builds every ray of every training image on the CPU, keeps four host tensors and, per iteration,
draws ``np.random.choice`` indices, gathers on the host and copies the batch to the device.  Here the rays are
generated on the device (csrc/raygen.cu), stored as packed 48-byte rows (csrc/raystore.cu) and a batch is one
gather kernel over indices drawn on the device -- no host work and no host->device copy in the training loop.
"""
import ctypes

import torch

from . import _lib
from .rays import ray_bundle_cuda


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class DeviceRayStore:
    """Same constructor arguments and ``get_training_rays_for_next_iter`` as ``TrainDataset``; ``images`` [n,H,W,3]
    and ``poses`` [n,4,4] may live on any device."""

    ROW = 12

    def __init__(self, poses, images, focal, ndc_rays=False, single_image_mode=False, device="cuda"):
        lib = _lib.load()
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("ddnerf_b200: DeviceRayStore needs a CUDA device (the path has no CPU fallback)")
        self.ndc = bool(ndc_rays)
        self.n_images, self.H, self.W = int(images.shape[0]), int(images.shape[1]), int(images.shape[2])
        self.focal, self.near, self.single_image_mode = focal, 1, bool(single_image_mode)
        self.rays_per_image = self.H * self.W
        self.total = self.n_images * self.rays_per_image
        with torch.cuda.device(self.dev):
            self.rows = torch.empty(self.total, self.ROW, device=self.dev, dtype=torch.float32)
            self._bad = torch.zeros(1, device=self.dev, dtype=torch.int32)
            for i in range(self.n_images):
                ro, rd, rad = ray_bundle_cuda(self.H, self.W, focal, poses[i], self.dev, ndc_near=self.near if self.ndc else None)
                tgt = torch.as_tensor(images[i]).to(self.dev, torch.float32).reshape(-1, 3).contiguous()
                dst = self.rows[i * self.rays_per_image:(i + 1) * self.rays_per_image]
                _lib.check(lib.ddnerf_raystore_pack(_p(ro), _p(rd), _p(rad), _p(tgt), self.rays_per_image, _p(dst), _stream()),
                           "raystore_pack")
        print(f"training set init finnished, {self.total} rays in the dataset")          # dataset.py:48

    def __len__(self):
        return self.total

    def get_training_rays_for_next_iter(self, number_of_rays, device=None, idxs=None, img_idx=None):
        """dataset.py:50-59.  Returns (origins [n,3], directions [n,3], radii [n,1], target [n,3]) on the store's
        device.  ``idxs`` (and ``img_idx`` in single-image mode) may be given -- e.g. the reference's own
        ``np.random.choice`` draws -- otherwise they are drawn on the device (with replacement, like the reference)."""
        lib = _lib.load()
        if device is not None and torch.device(device).type != "cuda":
            raise RuntimeError("ddnerf_b200: the ray store serves CUDA batches only")
        n = int(number_of_rays)
        with torch.cuda.device(self.dev):
            base, span = 0, self.total
            if self.single_image_mode:                                   # one random image per iteration (:55-58)
                if img_idx is None:
                    img_idx = int(torch.randint(0, self.n_images, (1,)).item())
                base, span = int(img_idx) * self.rays_per_image, self.rays_per_image
            if idxs is None:
                idxs = torch.randint(0, span, (n,), device=self.dev, dtype=torch.int64)
            else:
                idxs = torch.as_tensor(idxs).to(self.dev, torch.int64).contiguous()
                n = idxs.numel()
            ro = torch.empty(n, 3, device=self.dev)
            rd = torch.empty(n, 3, device=self.dev)
            rad = torch.empty(n, 1, device=self.dev)
            tgt = torch.empty(n, 3, device=self.dev)
            # rows of the chosen image only: the kernel bounds-checks against [0, total) after adding `base`
            _lib.check(lib.ddnerf_raystore_gather(_p(self.rows), self.total, _p(idxs), n, base, _p(ro), _p(rd), _p(rad), _p(tgt),
                                                  _p(self._bad), _stream()), "raystore_gather")
        return ro, rd, rad, tgt

    def check_indices(self):
        """True if no gather so far saw an out-of-range index (one device->host read; not for the hot loop)."""
        return int(self._bad.item()) == 0

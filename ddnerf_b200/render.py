"""The render loop of the reference (render_video.py:62-101, eval_nerf.py) on the device, frame by frame:

    camera pose -> rays (csrc/raygen.cu) -> model.run_iter(mode="validation") -> 8-bit images (csrc/frame.cu)

Rows f1 and f4 of SURVEY.md section 8f around the per-ray path.  ``FrameRenderer`` owns the static buffers of one
frame size and, with ``use_graph=True``, replays the whole frame as ONE CUDA graph (with several ranks the graph contains
the one NCCL all-reduce that makes the disparity range a whole-frame range): the only host->device traffic
of a frame is its 48-byte pose, the only device->host traffic the 8-bit images (4 bytes per pixel, 6 with the
side-by-side video frame) instead of 28 bytes of rays in and 16 bytes of float images out per pixel.
"""
import ctypes

import torch

from . import _lib
from .trainer import shard_rows


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def frame_to_u8(rgb, disp, want_video=False, minmax=None, workspace=None, out=None):
    """cast_to_image / cast_to_disparity_image (visualization.py:11-27) and, optionally, the [rgb | disparity] BGR
    video frame of render_video.py:96-101, for a float frame on the device.  rgb [H,W,3], disp [H,W] ->
    (rgb8 [H,W,3], disp8 [H,W], video [H,2W,3] | None) uint8 CUDA tensors.  ``minmax`` (2 floats on the device)
    overrides the disparity range, e.g. the range of the whole frame when ranks render row blocks."""
    lib = _lib.load()
    if not (rgb.is_cuda and disp.is_cuda):
        raise RuntimeError("ddnerf_b200: frame_to_u8 needs CUDA tensors (the path has no CPU fallback)")
    rgb, disp = rgb.contiguous().float(), disp.contiguous().float()
    H, W = disp.shape
    dev = rgb.device
    if minmax is None:
        minmax = torch.empty(2, device=dev)
        workspace = torch.zeros(4, device=dev, dtype=torch.int32) if workspace is None else workspace
        _lib.check(lib.ddnerf_frame_minmax(_p(disp), disp.numel(), _p(workspace), _p(minmax), _stream()), "frame_minmax")
    if out is None:
        out = (torch.empty(H, W, 3, device=dev, dtype=torch.uint8), torch.empty(H, W, device=dev, dtype=torch.uint8),
               torch.empty(H, 2 * W, 3, device=dev, dtype=torch.uint8) if want_video else None)
    _lib.check(lib.ddnerf_frame_pack_u8(_p(rgb), _p(disp), _p(minmax), _p(out[0]), _p(out[1]), _p(out[2]), H, W, _stream()),
               "frame_pack_u8")
    return out


class FrameRenderer:
    """Renders pixel rows [lo, hi) of H x W frames of ``model`` from camera poses.

    ``render(pose)`` returns (rgb8 [rows,W,3], disp8 [rows,W], video [rows,2W,3] | None) as pinned HOST uint8
    tensors (valid until the next call).  ``ndc_near``: forward-facing scenes (dataset_helpers.py:3-42)."""

    def __init__(self, model, height, width, focal, ndc_near=None, rank=0, world=1, want_video=False, use_graph=True):
        self.model, self.H, self.W, self.focal, self.ndc_near = model, int(height), int(width), float(focal), ndc_near
        self.lo, self.hi = shard_rows(self.H, rank, world)
        self.world = world
        self.dev = next(model.coarse.parameters()).device
        rows = self.hi - self.lo
        with torch.cuda.device(self.dev):
            self.pose_host = torch.zeros(12, dtype=torch.float32).pin_memory()
            self.pose_dev = torch.zeros(12, device=self.dev)
            self.rays = (torch.empty(rows, self.W, 3, device=self.dev), torch.empty(rows, self.W, 3, device=self.dev),
                         torch.empty(rows, self.W, 1, device=self.dev))
            self.minmax = torch.empty(2, device=self.dev)
            self.ws = torch.zeros(4, device=self.dev, dtype=torch.int32)
            self.out_dev = (torch.empty(rows, self.W, 3, device=self.dev, dtype=torch.uint8),
                            torch.empty(rows, self.W, device=self.dev, dtype=torch.uint8),
                            torch.empty(rows, 2 * self.W, 3, device=self.dev, dtype=torch.uint8) if want_video else None)
            self.out_host = tuple(None if t is None else torch.empty(t.shape, dtype=torch.uint8).pin_memory()
                                  for t in self.out_dev)
        # the masked mus/sigmas diagnostics of models.py:292-300 have data-dependent shapes (a host sync per pass);
        # the render loop never reads them
        model.record_distributions = False
        self.use_graph = bool(use_graph)
        self.capture_collective = True   # several ranks: capture the NCCL all-reduce into the frame's graph (one graph per frame)
        self._graph = None           # the whole frame (or, when the collective cannot be captured, the part before it)
        self._graph_tail = None      # several ranks: the part after it
        self._calls = 0
        self.float_out = None
        self.empty = self.hi <= self.lo      # more ranks than pixel rows: this rank only takes part in the collective

    def _head(self):
        """pose -> rays -> model -> rank-local disparity range."""
        lib = _lib.load()
        self.pose_dev.copy_(self.pose_host, non_blocking=True)
        if self.empty:
            self.minmax.fill_(float("-inf"))                 # neutral element of the MAX all-reduce of (-min, max)
            return
        _lib.check(lib.ddnerf_ray_bundle_dev(self.H, self.W, self.focal, _p(self.pose_dev), int(self.ndc_near is not None),
                                             float(self.ndc_near or 0.0), self.lo, self.hi, _p(self.rays[0]),
                                             _p(self.rays[1]), _p(self.rays[2]), _stream()), "ray_bundle_dev")
        with torch.no_grad():
            out = self.model.run_iter(*self.rays, mode="validation")
        rgb, disp = out[1]["rgb"], out[1]["disp"].contiguous()
        self.float_out = (rgb, disp)
        _lib.check(lib.ddnerf_frame_minmax(_p(disp), disp.numel(), _p(self.ws), _p(self.minmax), _stream()), "frame_minmax")
        if self.world > 1:
            self.minmax[0:1].neg_()                          # one MAX all-reduce of (-min, max) serves both ends

    def _collective(self):
        """Whole-frame disparity range when ranks render row blocks (the only collective of the render loop)."""
        import torch.distributed as dist
        dist.all_reduce(self.minmax, op=dist.ReduceOp.MAX)

    def _tail(self):
        """8-bit conversion and the copies to the pinned host images."""
        if self.empty:
            return
        if self.world > 1:
            self.minmax[0:1].neg_()
        frame_to_u8(self.float_out[0], self.float_out[1], minmax=self.minmax, out=self.out_dev)
        for dst, src in zip(self.out_host, self.out_dev):
            if dst is not None:
                dst.copy_(src, non_blocking=True)

    def _body(self):
        self._head()
        if self.world > 1:
            self._collective()
        self._tail()

    def render(self, pose):
        """pose: 4x4 (or 3x4) camera-to-world, any device.  Synchronises the stream before returning."""
        self.pose_host.copy_(torch.as_tensor(pose, dtype=torch.float32).detach().cpu().reshape(-1)[:12])
        self.model.eval()
        if self.use_graph and self._calls >= 2:
            if self._graph is None:
                # the packed bf16 weight images are refreshed by a host-side decision (TcState.refresh): capture with the
                # state dirty so the pack kernels are nodes of the graph and every replay renders the CURRENT weights
                # (a Trainer step or load_state_dict between frames would otherwise leave the graph on stale images)
                for net in {id(self.model.coarse): self.model.coarse, id(self.model.fine): self.model.fine}.values():
                    st = getattr(net, "_tc_state", None)
                    if st is not None:
                        st.dirty = True
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                one_graph = self.world == 1 or self.capture_collective
                try:
                    with torch.cuda.graph(g):
                        if one_graph:
                            self._body()                     # several ranks: the NCCL MAX all-reduce is a node of the graph
                        else:
                            self._head()
                except Exception:
                    if not (self.world > 1 and one_graph):
                        raise
                    # this NCCL / driver combination refuses to capture the collective: two graphs around an eager call
                    self.capture_collective, one_graph = False, False
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._head()
                self._graph = g
                if not one_graph:                            # the NCCL call stays eager, between two graphs
                    t = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(t, pool=g.pool()):
                        self._tail()
                    self._graph_tail = t
            self._graph.replay()
            if self._graph_tail is not None:
                self._collective()
                self._graph_tail.replay()
        else:
            self._body()
        self._calls += 1
        torch.cuda.current_stream().synchronize()
        return self.out_host

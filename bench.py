#!/usr/bin/env python
"""Benchmark of the DDNeRF / mip-NeRF per-ray hot path (BASELINE.json metric: train rays/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
                    [--workload cfg2|cfg1|cfg4] [--mlp-mode bf16|fp32]

One "step" = one training iteration of train_model.py:135-177 (run_iter on a ray batch, loss,
backward, Adam) on synthetic dataset-shaped rays and seeded random-init weights.  N=1 runs the
configuration BASELINE.json quotes for one B200: config_blender_mipnerf.yml, 4096 rays,
128+128 samples (cfg2).  N>1 (torchrun, one rank per GPU): same per-GPU batch, rays sharded by
rank, one NCCL all-reduce of the flat gradient bucket per step (weak scaling).

`--impl reference` times the UNMODIFIED reference (its stock `models.models.<Model>.run_iter`, the loss of
train_model.py:156-167, `loss.backward()` and its two `torch.optim.Adam` steps) on the host cores: the reference is
pure Python, `__graft_entry__.build()` copies its tree as is to baseline/_ref/ (git-ignored, shipped to the GPU box).
Every step is the full batch of the workload when the run fits the time budget, else a bounded ray sample (stated in
`cpu_baseline.sample`).  If baseline/_ref is missing the CPU restatement (oracle/ddnerf_oracle.py, pinned to the
reference by tests/test_oracle_golden.py) stands in, `kind: "port"`.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {
    # name: (preset, overrides, rays per GPU, description)
    "cfg2": ("config_blender_mipnerf", dict(num_coarse=128, num_fine=128), 4096,
             "config_blender_mipnerf.yml mip-NeRF IPE coarse/fine 128+128 samples, 4096-ray batch train step"),
    "cfg1": ("config_blender", dict(), 1024, "config_blender.yml DDNeRF 32+32 samples, 1024-ray batch train step"),
    "cfg4": ("config_360", dict(), 16384, "config_360.yml DDNeRF 32+32 samples, 16384 rays/GPU train step"),
}
MACS_TRAIN = {4: 610304 * 2 + 557696, 6: 610560 * 2 + 557696 + 2 * 128}   # fwd + dW + dX MACs per sample row


def mlp_flops_per_step(cfg, n_rays):
    """Algorithmic MLP FLOPs of one train step (3.557 MFLOP/sample, both passes)."""
    s0, s1 = cfg.nerf.train.num_coarse, cfg.nerf.train.num_fine
    is_dd = cfg.nerf.type == "DDNerfModel"
    return 2.0 * n_rays * (s0 * MACS_TRAIN[6 if is_dd else 4] + s1 * MACS_TRAIN[4])


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm, mx, reasons = [], 0.0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_batches(kind, n_rays, count, rank, near, far):
    """`count` host batches (pinned) of dataset-shaped rays + random targets."""
    from ddnerf_b200.rays import synth_rays
    batches = []
    for b in range(count):
        ro, rd, rad, _, _ = synth_rays(kind, n_rays, seed=97 * rank + b)
        tgt = torch.rand(n_rays, 3, generator=torch.Generator().manual_seed(5000 + 97 * rank + b))
        batches.append(tuple(t.pin_memory() if torch.cuda.is_available() else t for t in (ro, rd, rad, tgt)))
    return batches


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def cpu_baseline(cfg, kind, n_rays, steps, warmup):
    """The oracle's train step (fwd + bwd) on host cores; rays/s."""
    from oracle import ddnerf_oracle as orc
    from ddnerf_b200.rays import synth_rays
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    is_dd = cfg.nerf.type == "DDNerfModel"
    tp = cfg.train_params
    ocfg = orc.PathConfig(model=cfg.nerf.type, near=cfg.dataset.near, far=cfg.dataset.far,
                          num_coarse=cfg.nerf.train.num_coarse, num_fine=cfg.nerf.train.num_fine, perturb=True,
                          noise_std=cfg.nerf.train.radiance_field_noise_std, blender=cfg.dataset.type.lower() == "blender",
                          pdf_padding=tp.pdf_padding, gaussian_smooth_factor=tp.gaussian_smooth_factor,
                          dist_reg_coeficient=tp.dist_reg_coeficient, loss_coeficients=tp.loss_coeficients,
                          dp_coeficient=tp.dp_coeficient)
    pc = orc.init_mlp_params(is_dd, seed=42)
    pf = orc.init_mlp_params(False, seed=43) if is_dd else None
    ro, rd, rad, _, _ = synth_rays(kind, n_rays, seed=1)
    rays = orc.pack_rays(ro, rd, rad, ocfg.near, ocfg.far)
    g = torch.Generator().manual_seed(0)
    tgt = torch.rand(n_rays, 3, generator=g)
    s0, s1 = ocfg.num_coarse, ocfg.num_fine
    times = []
    for it in range(warmup + steps):
        rnd = dict(t_rand=torch.rand(n_rays, s0 + 1, generator=g), noise0=torch.randn(n_rays, s0, generator=g),
                   u_rand=torch.rand(n_rays, s1 + 1, generator=g), noise1=torch.randn(n_rays, s1, generator=g))
        t0 = time.perf_counter()
        orc.train_step(ocfg, pc, pf, rays, tgt, rnd)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return n_rays / sec, sec, threads


class ReferenceStep:
    """One training iteration of the reference's own loop (train_model.py:152-177) on the CPU, stock code path:
    `getattr(models, cfg.nerf.type)(cfg)` built from the reference's shipped YAML (plus the BASELINE.json workload
    overrides), `run_iter(mode="train")`, the loss of :156-167, `loss.backward()`, one Adam step per network."""

    def __init__(self, workload):
        from oracle import reference_loader as RL
        from ddnerf_b200.config import preset
        from ddnerf_b200.rays import synth_rays
        pname, over, n_rays, _ = WORKLOADS[workload]
        ours, kind = preset(pname, **over)                         # near/far/dist_reg as the native arm sets them
        root, ref_models, _ = RL.import_reference()
        cfg = RL.load_reference_cfg(pname)
        for mode in ("train", "validation"):
            cfg.nerf[mode].num_coarse, cfg.nerf[mode].num_fine = ours.nerf.train.num_coarse, ours.nerf.train.num_fine
        cfg.dataset.near, cfg.dataset.far = float(ours.dataset.near), float(ours.dataset.far)
        cfg.dataset.normalize_poses = False                        # (near/far above are already normalised, data_utils.py:67-74)
        cfg.train_params.dist_reg_coeficient = ours.train_params.dist_reg_coeficient      # train_model.py:124-125
        self.cfg, self.kind, self.n_full = cfg, kind, n_rays
        self.threads = os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        torch.manual_seed(cfg.experiment.randomseed)
        self.model = getattr(ref_models, cfg.nerf.type)(cfg)
        self.model.to("cpu")
        self.optims = [torch.optim.Adam(self.model.coarse.parameters(), lr=cfg.optimizer.lr)]
        if cfg.nerf.type != "GeneralMipNerfModel":                 # train_model.py:93-98
            self.optims.append(torch.optim.Adam(self.model.fine.parameters(), lr=cfg.optimizer.lr))
        self.synth = synth_rays
        self.root = root
        self._batches = {}

    def batch(self, n, b):
        if (n, b) not in self._batches:
            ro, rd, rad, _, _ = self.synth(self.kind, n, seed=b)
            tgt = torch.rand(n, 3, generator=torch.Generator().manual_seed(5000 + b))
            self._batches[(n, b)] = (ro, rd, rad, tgt)
        return self._batches[(n, b)]

    def step(self, n, b=0):
        ro, rd, rad, tgt = self.batch(n, b % 4)
        cfg = self.cfg
        self.model.train()
        t0 = time.perf_counter()
        out = self.model.run_iter(ro, rd, rad, mode="train", rgb_target=tgt)
        loss = 0.0
        for j in range(len(out)):
            loss = loss + cfg.train_params.loss_coeficients[j] * torch.nn.functional.mse_loss(out[j]["rgb"], tgt)
        if cfg.nerf.type == "DDNerfModel":
            loss = loss + cfg.train_params.dp_coeficient * out[1]["dp_loss"].mean()
        loss.backward()
        loss.item()
        for o in self.optims:
            o.step()
            o.zero_grad()
        return time.perf_counter() - t0


def reference_available():
    from oracle import reference_loader as RL
    return RL.reference_root() is not None


def bench_config(desc, n_rays):
    """The `config` object of the JSON line -- the same in both arms."""
    return {"workload": desc, "rays_per_gpu": n_rays,
            "l2": "per-step working set (activations + workspaces, GBs) far exceeds the 126 MB L2; no flush needed"}


def run_reference(args, cfg, kind, desc):
    """The reference arm: rank 0 alone, the reference's own CPU implementation on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_full = WORKLOADS[args.workload][2]
    K, W = args.steps, args.warmup
    import warnings
    warnings.filterwarnings("ignore")
    if reference_available():
        ref = ReferenceStep(args.workload)
        # size the run: one full-batch step is timed first (it is also a warm-up step); if K + W full batches do not fit
        # the time budget, every step is a bounded ray sample instead
        t_first = ref.step(n_full, 0)
        n = n_full
        if t_first * (K + max(W - 1, 0)) > args.ref_budget_s:
            n = int(n_full * args.ref_budget_s / (t_first * (K + max(W - 1, 0))))
            n = max(128, min(n_full, n // 128 * 128))
        for w in range(max(W - 1, 0)):
            ref.step(n, w + 1)
        times = [ref.step(n, s) for s in range(K)]
        kind_, threads = "reference", ref.threads
        sample = (f"{n}-ray batch per step ({'the full batch of the workload' if n == n_full else 'bounded sample of the ' + str(n_full) + '-ray batch'}), "
                  f"stock run_iter + loss + backward + Adam of the unmodified reference (baseline/_ref), torch {torch.__version__} CPU fp32")
    else:
        n = min(args.cpu_rays, n_full)
        rps_, sec_, threads = cpu_baseline(cfg, kind, n, K, max(1, min(W, 1)))
        times, kind_ = [sec_] * K, "port"
        sample = f"{n}-ray batch of the workload per step, fwd+bwd, torch CPU fp32 oracle (baseline/_ref not present)"
    sec = sum(times) / len(times)
    rps = n / sec
    line = {"impl": "reference", "metric": "train_rays_per_sec", "value": rps, "unit": "rays/s", "n_gpus": args.gpus,
            "steps": K, "warmup": W, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(desc, n_full),
            "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": threads, "kind": kind_, "sample": sample},
            "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "spread": {"ms_per_step_min": min(times) * 1e3, "ms_per_step_max": max(times) * 1e3}}
    emit(line)


def run_render(args, dev, world, rank, dist):
    """Full-frame render of BASELINE.json configs[2]: config_ff.yml DDNeRF, 1008x756 forward-facing NDC rays,
    16+16 samples, validation mode exactly as render_video.py:36-48,73-78 sets it up (pdf_padding off,
    gaussian_smooth_factor = final_smooth, det sampling, noise std 1.0), pixel rows split across ranks, no
    collective.  Returns the dict that goes under "render" in the JSON line."""
    from ddnerf_b200.config import preset
    from ddnerf_b200.models import models as M
    from ddnerf_b200.rays import frame as frame_preset, full_frame_rays
    from ddnerf_b200.trainer import shard_rows
    cfg, kind = preset("config_ff")
    cfg.train_params.pdf_padding = False
    cfg.train_params.gaussian_smooth_factor = cfg.train_params.final_smooth
    ro, rd, rad, near, far = full_frame_rays(kind)
    H, W = ro.shape[:2]
    lo, hi = shard_rows(H, rank, world)
    cfg.nerf.validation.chunksize = args.render_chunk if args.render_chunk > 0 else (hi - lo) * W
    host = tuple(t[lo:hi].contiguous().pin_memory() for t in (ro, rd, rad))
    torch.manual_seed(cfg.experiment.randomseed)
    model = M.DDNerfModel(cfg)
    model.to(dev)
    model.eval()
    model.record_distributions = False
    for net in (model.coarse, model.fine):
        net.mlp_mode = args.mlp_mode
    resident = tuple(t.to(dev) for t in host)
    img_host = torch.empty((hi - lo), W, 4, pin_memory=True)
    stage = [torch.empty_like(t, device=dev) for t in host]

    def frame_resident():
        with torch.no_grad():
            return model.run_iter(*resident, mode="validation")

    def frame_e2e():
        for dst, src in zip(stage, host):
            dst.copy_(src, non_blocking=True)
        with torch.no_grad():
            out = model.run_iter(*stage, mode="validation")
        img_host[..., :3].copy_(out[1]["rgb"], non_blocking=True)
        img_host[..., 3].copy_(out[1]["disp"], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    fH, fW, focal, c2w, _, _, _ = frame_preset(kind)
    from ddnerf_b200.render import FrameRenderer
    renderer = FrameRenderer(model, fH, fW, focal, ndc_near=1, rank=rank, world=world, want_video=False, use_graph=True)

    def frame_from_pose():
        # the render loop of render_video.py:62-101 per frame: pose -> rays (get_ray_bundle + ndc_mipnerf_rays: one device
        # kernel, csrc/raygen.cu) -> model -> 8-bit colour and disparity images (csrc/frame.cu) -> host; one replayed
        # CUDA graph per frame on one GPU.  Host->device traffic: the pose.
        renderer.render(c2w)

    def timed(fn, count):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(count):
            fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    frames = args.render_frames
    for _ in range(3):
        frame_resident()
    ms = timed(frame_resident, frames)
    frame_e2e()
    ms_e2e = timed(frame_e2e, frames)
    for _ in range(3):                                    # two eager frames, then the capture
        frame_from_pose()
    ms_pose = timed(frame_from_pose, frames)
    rays = H * W
    return {"metric": "render_rays_per_sec", "value": rays * frames / (ms * 1e-3), "unit": "rays/s",
            "ms_per_frame": ms / frames, "scaling": "strong",
            "e2e": {"value": rays * frames / (ms_pose * 1e-3), "unit": "rays/s", "ms_per_frame": ms_pose / frames,
                    "input": "camera pose in, 8-bit colour + disparity images out (ddnerf_b200/render.py: rays and image "
                             "conversion on the device, the frame replayed as one CUDA graph when n_gpus == 1)",
                    "h2d_bytes_per_frame": 48 * world, "d2h_bytes_per_frame": (hi - lo) * W * 4 * world},
            "e2e_host_rays": {"value": rays * frames / (ms_e2e * 1e-3), "unit": "rays/s", "ms_per_frame": ms_e2e / frames,
                              "input": "rays built on the host (the reference's loop), copied from pinned memory",
                              "h2d_bytes_per_frame": sum(t.numel() * 4 for t in host) * world,
                              "d2h_bytes_per_frame": img_host.numel() * 4 * world},
            "workload": "config_ff.yml DDNeRF 1008x756 full frame, 16+16 samples, validation mode, rows split over ranks",
            "frames": frames, "chunk_rays": int(cfg.nerf.validation.chunksize)}


def run_train_block(args, wl_name, dev, world, rank, dist, K, W):
    """A second training workload in the same line (weak scaling, rays per GPU fixed): BASELINE.json configs[3] --
    config_360.yml DDNeRF, 16,384 rays per GPU, 32 + 32 samples, gradient all-reduce of both networks' flat buckets --
    so that the scaling run covers the configuration BASELINE.json names for multi-GPU training.  Inputs resident in HBM,
    one CUDA graph per step, three blocks of K steps (median block reported)."""
    from ddnerf_b200.config import preset
    from ddnerf_b200.models import models as M
    from ddnerf_b200.trainer import Trainer
    pname, over, n_rays, desc = WORKLOADS[wl_name]
    cfg, kind = preset(pname, **over)
    torch.manual_seed(cfg.experiment.randomseed)
    model = getattr(M, cfg.nerf.type)(cfg)
    model.to(dev)
    for net in {id(model.coarse): model.coarse, id(model.fine): model.fine}.values():
        net.mlp_mode = args.mlp_mode
    trainer = Trainer(model, distributed=world > 1, use_graph=not args.no_graph)
    torch.manual_seed(4321 + rank)
    host = make_batches(kind, n_rays, 4, rank, cfg.dataset.near, cfg.dataset.far)
    resident = [tuple(t.to(dev) for t in b) for b in host]

    def block(count):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(count):
            trainer.step(*resident[s % len(resident)])
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    block(max(W, Trainer.GRAPH_WARMUP + 2))
    blocks = sorted(block(K) for _ in range(3))
    ms = blocks[1]
    flops = mlp_flops_per_step(cfg, n_rays)
    out = {"metric": "train_rays_per_sec", "value": n_rays * world * K / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms / K,
           "scaling": "weak", "workload": desc, "rays_per_gpu": n_rays, "steps": K,
           "spread_ms_per_step": [b / K for b in blocks],
           "mlp_tflops_over_whole_step": flops / (ms / K * 1e-3) / 1e12,
           "nccl_in_graph": bool(world > 1 and trainer.use_graph and trainer._graph_tail is None)}
    del trainer, model, resident
    torch.cuda.empty_cache()
    return out


_JSON_OUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line.  Libraries write there too (NCCL prints its version banner to fd 1 at
    communicator set-up whenever NCCL_DEBUG >= VERSION, from the environment or nccl.conf), so keep a private copy of
    the real stdout for the JSON line and point fd 1 at stderr for everything else."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    claim_stdout()
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def compositing_hbm(dev, n_rays_step, samples_step):
    """BASELINE.json's third figure, `compositing HBM GB/s`: the alpha-compositing kernels (csrc/composite.cu: forward and
    analytic backward of volume_render_radiance_field, volume_rendering_utils.py:6-84) timed alone, ALGORITHMIC bytes /
    CUDA-event time, against the measured copy bandwidth.  Two shapes: the sweep's 65,536 rays x 128 samples (configs[4],
    working set 235-370 MB > the 126 MB L2) and the benched step's own shape, where a launch moves 15-23 MB and runs for
    about 10 us (launch floor; inside the step these kernels are nodes of the step's CUDA graph).  Each kernel is captured
    into a one-node CUDA graph; an L2-sized buffer is rewritten between repetitions; median of 10."""
    from ddnerf_b200 import _lib, ops
    from ddnerf_b200.ops import _p, _stream
    lib = _lib.load()
    pk, pk_kind = peaks()
    flush = torch.zeros(64 * 1024 * 1024, device=dev)                 # 256 MB

    def timed_kernel(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep = fn()                                               # noqa: F841 (outputs live with the graph)
        g.replay()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in ev:
            flush.add_(1.0)
            a.record()
            g.replay()
            b.record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in ev)
        return ts[len(ts) // 2]

    out = {"kernel": "composite_fwd_kernel / composite_bwd_kernel (csrc/composite.cu)", "unit": "GB/s", "peak": pk["hbm_gbs"],
           "peak_kind": f"{pk_kind} copy bandwidth (MEASURED_PEAKS.json)",
           "bytes": "algorithmic: forward 28 B/sample + 40 B/ray, backward 44 B/sample + 24 B/ray", "shapes": []}
    for N, S in ((65536, 128), (n_rays_step, samples_step)):
        gen = torch.Generator(device=dev).manual_seed(3)
        raw = torch.randn(N, S, 4, device=dev, generator=gen)
        noise = torch.randn(N, S, device=dev, generator=gen)
        rd = torch.randn(N, 3, device=dev, generator=gen)
        t = torch.sort(torch.rand(N, S + 1, device=dev, generator=gen) * 4 + 2, dim=-1)[0]
        g_rgb, g_w = torch.randn(N, 3, device=dev, generator=gen), torch.randn(N, S, device=dev, generator=gen)
        g_raw = torch.empty(N, S, 4, device=dev)

        def fwd():
            with torch.no_grad():
                return ops.composite(raw, t, rd, noise, 1.0, None, False, True, False)

        def bwd():
            _lib.check(lib.ddnerf_composite_backward(_p(raw), 4, _p(t), _p(rd), rd.stride(0), _p(noise), 1.0, None, 0, 1,
                                                     _p(g_rgb), None, None, _p(g_w), None, None, _p(g_raw), None, N, S,
                                                     _stream()), "composite_backward")
        R = N * S
        b_f, b_b = R * 28 + N * 40, R * 44 + N * 24
        ms_f, ms_b = timed_kernel(fwd), timed_kernel(bwd)
        out["shapes"].append({"rays": N, "samples": S, "fwd_ms": ms_f, "fwd_gbs": b_f / ms_f / 1e6, "fwd_frac": b_f / ms_f / 1e6 / pk["hbm_gbs"],
                              "bwd_ms": ms_b, "bwd_gbs": b_b / ms_b / 1e6, "bwd_frac": b_b / ms_b / 1e6 / pk["hbm_gbs"]})
        del raw, noise, g_raw, g_w, t
    out["achieved"] = out["shapes"][0]["fwd_gbs"]
    out["frac"] = out["shapes"][0]["fwd_frac"]
    return out


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--mlp-mode", default=os.environ.get("DDNERF_MLP_MODE", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--cpu-rays", type=int, default=512, help="rays per step of the bounded CPU-baseline sample")
    ap.add_argument("--ref-budget-s", type=float, default=420.0,
                    help="--impl reference: wall-clock budget of the whole run; full batches if they fit, else a bounded sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the training step eagerly instead of replaying one CUDA graph")
    ap.add_argument("--no-render", action="store_true", help="skip the full-frame render leg (configs[2])")
    ap.add_argument("--no-hbm-kernels", action="store_true", help="skip the compositing HBM GB/s block")
    ap.add_argument("--no-cfg4", action="store_true", help="skip the DDNeRF 16,384 rays/GPU training block (configs[3])")
    ap.add_argument("--render-frames", type=int, default=5)
    ap.add_argument("--render-chunk", type=int, default=0, help="rays per chunk of the render leg (0: one chunk per rank)")
    args = ap.parse_args()

    from ddnerf_b200.config import preset
    pname, over, n_rays, desc = WORKLOADS[args.workload]
    cfg, kind = preset(pname, **over)

    if args.impl == "reference":
        run_reference(args, cfg, kind, desc)
        return

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (native arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    os.environ["DDNERF_MLP_MODE"] = args.mlp_mode
    from ddnerf_b200 import _lib, ops
    from ddnerf_b200.models import models as M
    from ddnerf_b200.trainer import Trainer
    lib = _lib.load()

    torch.manual_seed(cfg.experiment.randomseed)                       # identical weights on every rank
    model = getattr(M, cfg.nerf.type)(cfg)
    model.to(dev)
    for net in {id(model.coarse): model.coarse, id(model.fine): model.fine}.values():
        net.mlp_mode = args.mlp_mode
    # one CUDA graph per training step (with several ranks: two graphs around the eagerly launched NCCL all-reduce)
    use_graph = not args.no_graph
    trainer = Trainer(model, distributed=world > 1, use_graph=use_graph)
    torch.manual_seed(1234 + rank)                                     # per-rank draws inside the path

    K, W = args.steps, max(args.warmup, 3)
    host = make_batches(kind, n_rays, min(K, 8), rank, cfg.dataset.near, cfg.dataset.far)
    resident = [tuple(t.to(dev) for t in b) for b in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(run_step, count):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(count):
            run_step(s)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # ---- kernel-resident arm: inputs already in HBM ------------------------------------------
    def step_resident(s):
        trainer.step(*resident[s % len(resident)])

    l_before = lib.ddnerf_launch_count()
    step_resident(0)                                                   # eager (graph warm-up): kernels of one step
    launches_per_step = lib.ddnerf_launch_count() - l_before
    try:
        for s in range(1, max(W, Trainer.GRAPH_WARMUP + 1)):           # the last of these captures the graph
            step_resident(s)
        torch.cuda.synchronize()
    except Exception as exc:                                           # capture refused: fall back to eager launches
        if not trainer.use_graph:
            raise
        print(f"bench: CUDA graph capture failed ({type(exc).__name__}: {exc}); running eagerly", file=sys.stderr)
        trainer.use_graph, trainer._graph = False, None
        torch.cuda.synchronize()
        for s in range(W):
            step_resident(s)
    sampler = ClockSampler(local)
    sampler.start()
    # three blocks of exactly K steps each; the line reports the median block and the spread of the three
    blocks = sorted(timed(step_resident, K) for _ in range(3))
    ms_total = blocks[1]
    launches = launches_per_step * K                                   # (graph replays do not pass through the counter)
    clocks = sampler.stop()
    # per-kernel CUDA-event timing of the MLP launches: events cannot be recorded inside a replayed graph, so the
    # same steps are run eagerly (same kernels, same sizes) for the kernel-level numbers
    K_ev = min(K, 20)
    graphed, trainer.use_graph = trainer.use_graph, False
    timed(step_resident, 3)                                            # eager warm-up (allocator, clocks)
    ops.MLP_TIMING = []
    timed(step_resident, K_ev)
    trainer.use_graph = graphed
    mlp_events, ops.MLP_TIMING = ops.MLP_TIMING, None
    mlp_calls = len(mlp_events)
    by_tag = {}
    for a, b, tag in mlp_events:
        by_tag.setdefault(tag, []).append(a.elapsed_time(b))
    # per kernel: the MEDIAN launch duration times its launches per step (eager launches leave the GPU idle between
    # kernels, and single launches land on clock dips that the back-to-back graph replay does not see)
    for tag in by_tag:
        med = statistics.median(by_tag[tag])
        by_tag[tag] = [med] * len(by_tag[tag])
    mlp_ms = sum(sum(v) for v in by_tag.values()) * (K / K_ev)                 # scaled to the K timed steps

    # ---- end-to-end arm: pinned host rays -> device every step, loss read back every step ----
    # The loop a driver runs around Trainer.step: every iteration uploads ITS batch from pinned host memory and reads ITS
    # loss back.  Double-buffered: the upload of batch s+1 runs on a copy stream while step s computes, and the host reads
    # the loss of step s-1 while step s runs (it never runs more than one step ahead), instead of the reference loop's
    # upload -> step -> .item() serialisation (train_model.py:152-172).
    stage = [[torch.empty_like(t, device=dev) for t in host[0]] for _ in range(2)]
    loss_host = torch.empty(2, 3, pin_memory=True)
    copy_stream = torch.cuda.Stream(device=dev)
    up_done = [torch.cuda.Event() for _ in range(2)]
    used = [torch.cuda.Event() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    e2e_state = {"pending": None, "losses": []}

    def upload(s):
        buf = s % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(used[buf])                          # the step that last read this buffer has finished
            for dst, src in zip(stage[buf], host[s % len(host)]):
                dst.copy_(src, non_blocking=True)
            up_done[buf].record(copy_stream)

    def e2e_begin():
        for ev in used:
            ev.record()
        e2e_state["pending"] = None
        upload(0)

    def step_e2e(s):
        buf = s % 2
        main = torch.cuda.current_stream()
        main.wait_event(up_done[buf])
        loss, mse = trainer.step(*stage[buf])
        used[buf].record(main)
        loss_host[buf, 0:1].copy_(loss.reshape(1), non_blocking=True)
        loss_host[buf, 1:3].copy_(mse, non_blocking=True)
        loss_ev[buf].record(main)
        upload(s + 1)                                                  # overlaps this step's kernels
        if e2e_state["pending"] is not None:                           # read the PREVIOUS step's loss (host <= 1 step ahead)
            pb = e2e_state["pending"]
            loss_ev[pb].synchronize()
            e2e_state["losses"].append(float(loss_host[pb, 0]))
        e2e_state["pending"] = buf

    def e2e_end():
        pb = e2e_state["pending"]
        loss_ev[pb].synchronize()
        e2e_state["losses"].append(float(loss_host[pb, 0]))
        copy_stream.synchronize()

    def e2e_block(count):
        e2e_begin()
        for s in range(count):
            step_e2e(s)
        e2e_end()

    e2e_block(2)
    ms_e2e = sorted(timed(lambda s: e2e_block(K) if s == 0 else None, 1) for _ in range(3))[1]
    assert len(e2e_state["losses"]) == 2 + 3 * K and all(x == x for x in e2e_state["losses"])     # every loss arrived, none NaN

    # ---- the same loop fed by the device-resident ray store (row f3): no host work, no host->device batch copy ----
    from ddnerf_b200.raystore import DeviceRayStore
    from ddnerf_b200.rays import frame as frame_preset
    sH, sW, sfocal, spose, _, _, sndc = frame_preset(kind)
    n_img = 4
    g_img = torch.Generator().manual_seed(77 + rank)
    store = DeviceRayStore(torch.stack([spose] * n_img), torch.rand(n_img, sH, sW, 3, generator=g_img), sfocal,
                           ndc_rays=sndc, device=dev)

    def step_store(s):
        loss, mse = trainer.step(*store.get_training_rays_for_next_iter(n_rays, dev))
        loss_host[0, 0:1].copy_(loss.reshape(1), non_blocking=True)
        loss_host[0, 1:3].copy_(mse, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for s in range(2):
        step_store(s)
    ms_store = timed(step_store, K)

    total_rays = n_rays * world
    value = total_rays * K / (ms_total * 1e-3)
    e2e = total_rays * K / (ms_e2e * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    pk, pk_kind = peaks()
    flops_step = mlp_flops_per_step(cfg, n_rays)
    tf_achieved = flops_step * K / (mlp_ms * 1e-3) / 1e12 if mlp_ms > 0 else None
    peak_tf = pk["bf16_tflops_sustained"]
    e2e_store = {"value": total_rays * K / (ms_store * 1e-3), "unit": "rays/s", "ms_per_step": ms_store / K,
                 "input": "batch drawn and gathered on the device from the resident ray store (ddnerf_b200/raystore.py)",
                 "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 12, "store_rays": len(store)}
    line = {
        "metric": "train_rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.mlp_mode == "bf16" else "f32", "data": "synthetic",
        "config": bench_config(desc, n_rays),
        "impl_detail": {"mlp_mode": args.mlp_mode, "cuda_graph": bool(trainer.use_graph),
                        "nccl_in_graph": bool(world > 1 and trainer.use_graph and trainer._graph_tail is None)},
        "spread": {"blocks": 3, "steps_per_block": K, "ms_per_step": [b / K for b in blocks],
                   "rel": (blocks[2] - blocks[0]) / blocks[1]},
        "e2e": {"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
                "ms_per_step": ms_e2e / K,
                "loop": "pinned-host batch uploaded every step (double-buffered on a copy stream), loss + both mse read back "
                        "every step (host waits for step s-1's loss while step s runs)"},
        "e2e_ray_store": e2e_store,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": tf_achieved, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": (tf_achieved / peak_tf) if tf_achieved else None, "traffic": None,
                     "kernel": "NeRF MLP fwd+bwd (K1), %d launches/step; median CUDA-event duration of each kernel over %d eagerly launched steps" % (mlp_calls // max(K_ev, 1), K_ev),
                     "peak_kind": f"{pk_kind} bf16 sustained (MEASURED_PEAKS.json)",
                     "mlp_ms_per_step": mlp_ms / K if K else None},
    }
    # per-kernel view of the MLP (CUDA events around each launch, averaged over the timed steps): the two chain
    # kernels against the tensor roofline, the weight-gradient kernel against HBM (its algorithmic bytes: the
    # saved tile images it has to read, 10.6 KB per sample row incl. the layer-5 / view-branch re-reads)
    rows_step = n_rays * (cfg.nerf.train.num_coarse + cfg.nerf.train.num_fine)
    algo = {"fwd": ("tensor", 2.0 * 610304 * rows_step, "TFLOP/s", peak_tf, 1e12),
            "dx": ("tensor", 2.0 * 557696 * rows_step, "TFLOP/s", peak_tf, 1e12),
            "dw": ("hbm", 1400.0 * 1024 / 128 * rows_step, "GB/s", pk["hbm_gbs"], 1e9)}
    breakdown = []
    for tag, (bound, work, unit, peak_v, scale) in algo.items():
        if tag in by_tag:
            ms_step = sum(by_tag[tag]) / K_ev
            ach = work / (ms_step * 1e-3) / scale
            chain = "mlp_tc_pair_kernel" if os.environ.get("DDNERF_TC_PAIR", "1") != "0" else "mlp_tc_chain_kernel"
            breakdown.append({"kernel": {"fwd": chain + "<0> (forward + activation saves)",
                                         "dx": chain + "<1> (dX chain)", "dw": "mlp_tc_dw_kernel (dW, db)"}[tag],
                              "bound": bound, "launches_per_step": len(by_tag[tag]) // K_ev, "ms_per_step": ms_step,
                              "achieved": ach, "peak": peak_v, "unit": unit, "frac": ach / peak_v})
            if bound == "hbm":
                # the denominator is the measured COPY bandwidth (half reads, half writes); this kernel only reads, and a
                # read-only stream runs faster than a copy (DRAM bus turnarounds), so the fraction can pass 1
                breakdown[-1]["peak_kind"] = "measured copy bandwidth (MEASURED_PEAKS.json); the kernel is a read-only stream"
                breakdown[-1]["frac_of_spec_8TBs"] = ach / 8000.0
    if breakdown:
        line["roofline"]["breakdown"] = breakdown
        # DRAM bytes per step of the three MLP kernels from the committed ncu --set full captures
        # (profiles/r02j_ncu_mlp_tc_{pair_fwd_rays,pair_dx,dw}.md: read + write per launch at 524,288 rows), scaled by rows
        per_row = ((0.1007e9 + 2.7837e9) + (0.1606e9 + 2.5575e9) + (5.8592e9 + 0.0064e9)) / 524288.0
        line["roofline"]["traffic"] = per_row * rows_step if args.mlp_mode == "bf16" else None
        line["roofline"]["traffic_note"] = ("dram__bytes_read+write of the forward (ray-fed, with saves), dX chain and dW kernels from "
                                            "profiles/r02j_ncu_mlp_tc_{pair_fwd_rays,pair_dx,dw}.md, per step")
    if not args.no_cfg4 and args.workload != "cfg4":
        line["train_cfg4"] = run_train_block(args, "cfg4", dev, world, rank, dist, K, W)
    if not args.no_render:
        line["render"] = run_render(args, dev, world, rank, dist)
    if rank == 0 and world == 1 and not args.no_hbm_kernels:
        line["compositing"] = compositing_hbm(dev, n_rays, cfg.nerf.train.num_coarse)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_cpu = min(args.cpu_rays, n_rays)
        rps, sec, threads = cpu_baseline(cfg, kind, n_cpu, 2, 1)
        port = {"value": rps, "unit": "rays/s", "cores": threads, "kind": "port",
                "sample": f"{n_cpu}-ray batch of the workload, fwd+bwd, 1 warm-up + 2 timed steps, torch CPU fp32 oracle "
                          "(searchsorted instead of the reference's O(S^2) mask search: faster than the reference itself)"}
        if reference_available():
            import warnings
            warnings.filterwarnings("ignore")
            ref = ReferenceStep(args.workload)
            ref.step(n_cpu, 0)
            secs = [ref.step(n_cpu, b) for b in (1, 2)]
            rsec = sum(secs) / len(secs)
            line["cpu_baseline"] = {"value": n_cpu / rsec, "unit": "rays/s", "cores": ref.threads, "kind": "reference",
                                    "sample": f"{n_cpu}-ray batch of the workload, 1 warm-up + 2 timed steps of the unmodified "
                                              "reference's run_iter + loss + backward + Adam (baseline/_ref), torch CPU fp32; "
                                              "`bench.py --impl reference` times full batches"}
            line["cpu_baseline_port"] = port
        else:
            line["cpu_baseline"] = port
    if rank == 0:
        emit(line)
    if world > 1:
        # Teardown: the CUDA graphs that captured NCCL kernels must be gone before the communicator is destroyed (destroying
        # it under live graphs can block for ever); a timer ends the process if the destroy still does not return.
        import gc
        import threading
        sys.stdout.flush()
        torch.cuda.synchronize()
        dist.barrier()
        del trainer, model
        gc.collect()
        torch.cuda.synchronize()
        killer = threading.Timer(20.0, lambda: os._exit(0))
        killer.daemon = True
        killer.start()
        dist.destroy_process_group()
        killer.cancel()


if __name__ == "__main__":
    main()

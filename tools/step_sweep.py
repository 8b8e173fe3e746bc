#!/usr/bin/env python
"""Training-step sweep of BASELINE.json configs[4]: ray batch 4K-64K rays/GPU x 64-256 samples per ray
(coarse + fine), DDNeRF vs mip-NeRF, one GPU.  Every point is a full training iteration (sampling, encoding,
both MLP passes forward + backward, compositing, losses, Adam) replayed as one CUDA graph, timed with CUDA
events after warm-up; reported as rays/s and as MLP throughput against the measured sustained bf16 peak
(algorithmic 3.557 MFLOP per sample row).

    python tools/step_sweep.py [--out profiles/r01b_step_sweep] [--steps 10] [--max-rows 2200000]

Points whose saved activations would not fit in one piece (rows per pass above --max-rows) run with the batch split
into ray chunks of cfg.nerf.train.chunksize and the backward of each chunk right after its forward (gradient
accumulation, Trainer.accumulate_chunks) -- the `chunks` column.
"""
import argparse
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def time_point(cfg, kind, n_rays, dev, steps):
    import bench
    from ddnerf_b200.models import models as M
    from ddnerf_b200.trainer import Trainer
    torch.manual_seed(cfg.experiment.randomseed)
    model = getattr(M, cfg.nerf.type)(cfg)
    model.to(dev)
    for net in (model.coarse, model.fine):
        net.mlp_mode = "bf16"
    trainer = Trainer(model, distributed=False, use_graph=True)
    batch = tuple(t.to(dev) for t in bench.make_batches(kind, n_rays, 1, 0, cfg.dataset.near, cfg.dataset.far)[0])
    try:
        for _ in range(4):
            trainer.step(*batch)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            trainer.step(*batch)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps
    finally:
        del trainer, model, batch
        torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/step_sweep")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--max-rows", type=int, default=4_300_000)
    args = ap.parse_args()
    import bench
    from ddnerf_b200.config import preset
    from ddnerf_b200.models import models as M
    from ddnerf_b200.trainer import Trainer
    peak = bench.peaks()[0]["bf16_tflops_sustained"]
    dev = torch.device("cuda:0")
    rows = []
    for pname in ("config_blender_mipnerf", "config_blender"):
        for n_rays in (4096, 16384, 65536):
            for s in (32, 64, 128):
                cfg, kind = preset(pname, num_coarse=s, num_fine=s)
                name = "DDNeRF" if cfg.nerf.type == "DDNerfModel" else "mip-NeRF"
                rec = dict(model=name, rays=n_rays, samples=2 * s)
                chunk = n_rays
                while chunk * s > args.max_rows:                 # saved activations of one chunk: 5.4 KB per row and pass
                    chunk //= 2
                rec["chunks"] = n_rays // chunk
                cfg.nerf.train.chunksize = chunk
                try:
                    ms = time_point(cfg, kind, n_rays, dev, args.steps)
                except torch.OutOfMemoryError:
                    torch.cuda.empty_cache()
                    rec["skipped"] = f"{n_rays * s} rows per pass: out of device memory"
                    rows.append(rec)
                    continue
                tf = bench.mlp_flops_per_step(cfg, n_rays) / (ms * 1e-3) / 1e12
                rec.update(ms_per_step=ms, rays_per_s=n_rays / (ms * 1e-3), mlp_tflops_whole_step=tf, frac_of_sustained_peak=tf / peak)
                rows.append(rec)
                print(rec, flush=True)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out + ".json", "w") as f:
        json.dump(dict(peak_tflops=peak, rows=rows), f, indent=1)
    with open(args.out + ".md", "w") as f:
        f.write("# Training-step sweep (BASELINE.json configs[4]), one B200, bf16 MLP, one CUDA graph per step\n\n")
        f.write(f"MLP column: algorithmic MLP FLOPs of the step / WHOLE step time, against {peak} TFLOP/s (measured sustained bf16)\n\n")
        f.write("| model | rays | samples/ray (coarse+fine) | ray chunks | ms/step | rays/s | MLP TFLOP/s over the whole step | of peak |\n|---|---:|---:|---:|---:|---:|---:|---:|\n")
        for r in rows:
            if "skipped" in r:
                f.write(f"| {r['model']} | {r['rays']} | {r['samples']} | {r.get('chunks', 1)} | — | — | — | skipped: {r['skipped']} |\n")
            else:
                f.write(f"| {r['model']} | {r['rays']} | {r['samples']} | {r.get('chunks', 1)} | {r['ms_per_step']:.3f} | {r['rays_per_s']:.0f} | "
                        f"{r['mlp_tflops_whole_step']:.0f} | {100 * r['frac_of_sustained_peak']:.1f} % |\n")


if __name__ == "__main__":
    main()

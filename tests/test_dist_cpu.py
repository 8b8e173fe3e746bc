"""CPU suite, world_size 2 over gloo: the data-parallel host logic of ddnerf_b200/trainer.py
(flat gradient bucket, the single all-reduce, the 1/world scaling, pixel-row sharding for rendering).
The per-rank gradients come from the CPU oracle (test infrastructure): two ranks that each take half
of a ray batch must end with the gradient of the whole batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ddnerf_oracle as orc
from ddnerf_b200.rays import synth_rays
from ddnerf_b200.trainer import FlatBucket, allreduce_gradients, shard_rows


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem(N):
    cfg = orc.PathConfig(model="GeneralMipNerfModel", near=2.0, far=6.0, num_coarse=8, num_fine=8, perturb=True,
                         noise_std=1.0, blender=True, pdf_padding=True, gaussian_smooth_factor=1.7,
                         dist_reg_coeficient=0.03, loss_coeficients=(1.0, 0.1), dp_coeficient=0.1)
    ro, rd, rad, near, far = synth_rays("blender", N, seed=5)
    rays = orc.pack_rays(ro, rd, rad, near, far)
    g = torch.Generator().manual_seed(9)
    target = torch.rand(N, 3, generator=g)
    rnd = dict(t_rand=torch.rand(N, 9, generator=g), noise0=torch.randn(N, 8, generator=g),
               u_rand=torch.rand(N, 9, generator=g), noise1=torch.randn(N, 8, generator=g))
    return cfg, rays, target, rnd


class _Net(torch.nn.Module):
    """parameter container with the reference's names (state_dict of MipNeRFModel)"""

    def __init__(self, params):
        super().__init__()
        self.names = list(params)
        for k, v in params.items():
            self.register_parameter(k.replace(".", "__"), torch.nn.Parameter(v.clone()))


def _worker(rank, world, port, N, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    try:
        cfg, rays, target, rnd = _problem(N)
        lo, hi = (N * rank) // world, (N * (rank + 1)) // world          # this rank's rays
        params = orc.init_mlp_params(False, seed=3)                        # replicated weights
        _, _, grads, _ = orc.train_step(cfg, params, None, rays[lo:hi], target[lo:hi], {k: v[lo:hi] for k, v in rnd.items()})
        net = _Net(params)
        bucket = FlatBucket(net)
        for (name, p) in zip(net.names, net.parameters()):
            p.grad = grads[name].clone()
        bucket.gather_grads()
        allreduce_gradients([bucket], world)
        if rank == 0:
            out.put((bucket.grad / world).clone())
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_equals_full_batch():
    N, world = 64, 2
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = out.get()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    cfg, rays, target, rnd = _problem(N)
    params = orc.init_mlp_params(False, seed=3)
    _, _, grads, _ = orc.train_step(cfg, params, None, rays, target, rnd)
    ref = torch.cat([grads[k].reshape(-1) for k in params])
    assert got.shape == ref.shape
    err = (got - ref).abs().max() / ref.abs().max()
    assert err < 1e-5, err


def test_flat_bucket_views_and_gather():
    params = orc.init_mlp_params(True, seed=1)
    net = _Net(params)
    before = {k: v.clone() for k, v in zip(net.names, net.parameters())}
    b = FlatBucket(net)
    assert b.flat.numel() == sum(v.numel() for v in params.values()) == 612998
    for k, p in zip(net.names, net.parameters()):
        assert torch.equal(p, before[k]) and p.data_ptr() >= b.flat.data_ptr()      # views into the bucket
        p.grad = torch.full_like(p, 2.0)
    b.gather_grads()
    assert torch.all(b.grad == 2.0) and all(p.grad is None for p in net.parameters())


@pytest.mark.parametrize("H,world", [(756, 1), (756, 2), (756, 8), (800, 3), (5, 8)])
def test_render_row_sharding(H, world):
    spans = [shard_rows(H, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == H
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 == b0 and a0 <= a1
    assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_flat_bucket_state_dict_has_torch_adam_layout():
    """FlatBucket.state_dict() is what torch.optim.Adam.state_dict() holds for the same parameters (the reference's
    optimizer_{1,2}_state_dict, train_model.py:110-118,249-258): it loads into a real Adam, and a real Adam's state loads back."""
    params = orc.init_mlp_params(False, seed=2)
    net = _Net(params)
    b = FlatBucket(net)
    g = torch.Generator().manual_seed(0)
    b.exp_avg.copy_(torch.randn(b.flat.shape, generator=g))
    b.exp_avg_sq.copy_(torch.rand(b.flat.shape, generator=g))
    b.step = 7
    sd = b.state_dict(lr=1e-4)
    ref = _Net(params)
    opt = torch.optim.Adam(ref.parameters(), lr=5e-4)
    opt.load_state_dict(sd)
    assert opt.param_groups[0]["lr"] == 1e-4 and opt.param_groups[0]["betas"] == (0.9, 0.999)
    off = 0
    for p in ref.parameters():
        st = opt.state[p]
        assert float(st["step"]) == 7.0
        assert torch.equal(st["exp_avg"].reshape(-1), b.exp_avg[off:off + p.numel()])
        assert torch.equal(st["exp_avg_sq"].reshape(-1), b.exp_avg_sq[off:off + p.numel()])
        off += p.numel()
    b2 = FlatBucket(_Net(params))
    b2.load_state_dict(opt.state_dict())
    assert b2.step == 7 and torch.equal(b2.exp_avg, b.exp_avg) and torch.equal(b2.exp_avg_sq, b.exp_avg_sq)
    # a fresh optimizer (no steps yet) round-trips to zero moments
    b3 = FlatBucket(_Net(params))
    b3.load_state_dict(torch.optim.Adam(_Net(params).parameters()).state_dict())
    assert b3.step == 0 and not b3.exp_avg.any()


def test_flat_bucket_detects_re_allocated_parameters():
    net = _Net(orc.init_mlp_params(False, seed=2))
    b = FlatBucket(net)
    b.check_alias()
    net.double()                                         # re-allocates every parameter (what module.to(other device) does)
    with pytest.raises(RuntimeError, match="no longer aliases"):
        b.check_alias()


def test_draw_randoms_host_logic():
    """GeneralMipNerfModel._draw_randoms: the four random tensors of one predict() (samplers.py:57,102/165,
    volume_rendering_utils.py:31) come from one rand + one randn call as contiguous [N, n] tensors; injected tensors pass
    through; modes that do not use a tensor get None (validation: perturb off)."""
    from ddnerf_b200.config import preset
    from ddnerf_b200.models import models as M
    cfg, _ = preset("config_blender", num_coarse=8, num_fine=12)
    model = M.DDNerfModel(cfg)
    N = 5
    r = model._draw_randoms(N, "train", torch.device("cpu"))
    assert r["t_rand"].shape == (N, 9) and r["u_rand"].shape == (N, 13)
    assert r["noise0"].shape == (N, 8) and r["noise1"].shape == (N, 12)
    assert all(v.is_contiguous() and v.dtype == torch.float32 for v in r.values())
    assert 0.0 <= float(r["t_rand"].min()) and float(r["u_rand"].max()) < 1.0
    # the two uniform tensors are the halves of one flat draw (same for the normal ones)
    assert r["u_rand"].data_ptr() == r["t_rand"].data_ptr() + N * 9 * 4
    assert r["noise1"].data_ptr() == r["noise0"].data_ptr() + N * 8 * 4
    v = model._draw_randoms(N, "validation", torch.device("cpu"))
    assert v["t_rand"] is None and v["u_rand"] is None                  # perturb = False
    assert v["noise0"].shape == (N, 8)                                  # radiance_field_noise_std stays 1.0
    inj = torch.full((N, 13), 0.25)
    model.randoms = {"u_rand": inj}
    r = model._draw_randoms(N, "train", torch.device("cpu"))
    assert r["u_rand"].data_ptr() == inj.data_ptr() and r["t_rand"].shape == (N, 9)
    cfg.nerf.train.radiance_field_noise_std = 0.0
    r = model._draw_randoms(N, "train", torch.device("cpu"))
    assert r["noise0"] is None and r["noise1"] is None

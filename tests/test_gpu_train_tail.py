"""Row f2 (the training-step tail, train_model.py:86-98, 146-177) against PyTorch itself: the fused Adam kernels against
``torch.optim.Adam`` on the same gradients, the fused photometric loss against ``F.mse_loss`` and its autograd, the
device-side schedule against the host formulas of the driver loop, and the graphed Trainer queued without host syncs
against the eager trajectory."""
import ctypes
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _net(depth_head, seed):
    from oracle import ddnerf_oracle as orc
    from ddnerf_b200.models import base_architectures as BA
    net = (BA.DepthMipNeRFModel if depth_head else BA.MipNeRFModel)(
        hidden_size=256, max_ipe_deg=16, num_encoding_fn_dir=4, include_input_xyz=False, include_input_dir=True)
    net.load_state_dict(orc.init_mlp_params(depth_head, seed=seed))
    return net


@pytest.mark.parametrize("variant", ["host_scalars", "device_scalars"])
def test_fused_adam_matches_torch_optim_adam(variant):
    """10 steps of ddnerf_adam_step / ddnerf_adam_step_dev on the flat bucket against torch.optim.Adam (defaults of
    train_model.py:86-98: lr from learning_rate_decay, betas (0.9, 0.999), eps 1e-8) fed the same gradients: parameters
    and both moments within 1e-6 relative."""
    from ddnerf_b200 import ops
    from ddnerf_b200.general_utils.nerf_helpers import learning_rate_decay
    from ddnerf_b200.trainer import FlatBucket
    torch.manual_seed(0)
    ref_net, net = _net(True, 3).to(DEV), _net(True, 3).to(DEV)
    opt = torch.optim.Adam(ref_net.parameters(), lr=0.0005)
    b = FlatBucket(net)
    g = torch.Generator(device=DEV).manual_seed(5)
    hyper = torch.zeros(10, device=DEV)
    for i in range(10):
        lr = learning_rate_decay(i, 0.0005, 5e-6, 200001, lr_delay_steps=2500, lr_delay_mult=0.01)
        for pg in opt.param_groups:
            pg["lr"] = lr
        off = 0
        for p in ref_net.parameters():
            # gradient scales spanning several decades, some exact zeros (eps matters there)
            gr = torch.randn(p.shape, device=DEV, generator=g) * 10.0 ** torch.randint(-6, 1, (1,), device=DEV, generator=g).item()
            gr[torch.rand(p.shape, device=DEV, generator=g) < 0.05] = 0.0
            p.grad = gr
            b.grad[off:off + p.numel()].copy_(gr.reshape(-1))
            off += p.numel()
        opt.step()
        if variant == "host_scalars":
            b.adam(lr)
        else:
            t = i + 1
            hyper.copy_(torch.tensor([lr, 0.9, 0.999, 1e-8, 1 - 0.9 ** t, math.sqrt(1 - 0.999 ** t), 1.0, 0.0, 1 - 0.9, 1 - 0.999]))
            b.adam_dev(hyper)
            b.step += 1
    torch.cuda.synchronize()
    ref = torch.cat([p.detach().reshape(-1) for p in ref_net.parameters()])
    rel = ((b.flat - ref).abs().max() / ref.abs().max()).item()
    assert rel < 1e-6, rel
    # every nn.Parameter of the module still aliases the bucket
    for p, q in zip(net.parameters(), ref_net.parameters()):
        assert torch.allclose(p, q, rtol=0, atol=1e-6 * ref.abs().max().item())
    sd = opt.state_dict()["state"]
    off = 0
    for i, p in enumerate(ref_net.parameters()):
        n = p.numel()
        for key, buf in (("exp_avg", b.exp_avg), ("exp_avg_sq", b.exp_avg_sq)):
            want = sd[i][key].reshape(-1)
            assert ((buf[off:off + n] - want).abs().max() <= 1e-6 * want.abs().max() + 1e-30)
        off += n


def test_adam_state_dict_interchanges_with_torch_optim_adam():
    """FlatBucket.state_dict() loads into torch.optim.Adam and back (the reference's checkpoints, train_model.py:110-118,
    249-258), and continuing from either side gives the same parameters."""
    from ddnerf_b200.trainer import FlatBucket
    net, ref_net = _net(False, 4).to(DEV), _net(False, 4).to(DEV)
    b = FlatBucket(net)
    g = torch.Generator(device=DEV).manual_seed(1)
    for _ in range(3):
        b.grad.copy_(torch.randn(b.grad.shape, device=DEV, generator=g) * 1e-3)
        b.adam(3e-4)
    ref_net.load_state_dict(net.state_dict())
    opt = torch.optim.Adam(ref_net.parameters(), lr=3e-4)
    opt.load_state_dict(b.state_dict(lr=3e-4))
    gr = torch.randn(b.grad.shape, device=DEV, generator=g) * 1e-3
    off = 0
    for p in ref_net.parameters():
        p.grad = gr[off:off + p.numel()].view(p.shape).clone()
        off += p.numel()
    opt.step()
    b.grad.copy_(gr)
    b.adam(3e-4)
    ref = torch.cat([p.detach().reshape(-1) for p in ref_net.parameters()])
    assert ((b.flat - ref).abs().max() / ref.abs().max()).item() < 1e-6
    # and the other direction
    b2 = FlatBucket(_net(False, 4).to(DEV))
    b2.load_state_dict(opt.state_dict())
    assert b2.step == 4
    assert (b2.exp_avg - b.exp_avg).abs().max().item() <= 1e-6 * b.exp_avg.abs().max().item()
    assert (b2.exp_avg_sq - b.exp_avg_sq).abs().max().item() <= 1e-6 * b.exp_avg_sq.abs().max().item()


@pytest.mark.parametrize("N", [1, 37, 4096, 100003])
def test_fused_mse_matches_torch(N):
    """ddnerf_mse_loss against F.mse_loss (train_model.py:159-162) and its autograd: both losses within 1e-6 relative,
    cotangents of coef0*mse0 + coef1*mse1 within 1e-6 of autograd's."""
    from ddnerf_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(N)
    rgb0 = torch.rand(N, 3, device=DEV, generator=g).requires_grad_(True)
    rgb1 = torch.rand(N, 3, device=DEV, generator=g).requires_grad_(True)
    tgt = torch.rand(N, 3, device=DEV, generator=g)
    c0, c1 = 1.0, 0.1
    m0, m1 = torch.nn.functional.mse_loss(rgb0, tgt), torch.nn.functional.mse_loss(rgb1, tgt)
    (c0 * m0 + c1 * m1).backward()
    mse, g0, g1 = ops.mse_loss_and_grad(rgb0, rgb1, tgt, c0, c1)
    assert abs(mse[0].item() - m0.item()) <= 2e-6 * m0.item()
    assert abs(mse[1].item() - m1.item()) <= 2e-6 * m1.item()
    assert mse.shape == (3,) and abs(mse[2].item() - (c0 * m0 + c1 * m1).item()) <= 2e-6 * (c0 * m0 + c1 * m1).item()
    assert (g0 - rgb0.grad).abs().max().item() <= 1e-6 * rgb0.grad.abs().max().item() + 1e-12
    assert (g1 - rgb1.grad).abs().max().item() <= 1e-6 * rgb1.grad.abs().max().item() + 1e-12
    # single-output form (mip-NeRF validation path)
    mse_a, ga, none = ops.mse_loss_and_grad(rgb0, None, tgt, c0, 0.0)
    assert none is None and abs(mse_a[0].item() - m0.item()) <= 2e-6 * m0.item()


def test_device_schedule_matches_driver_formulas():
    """ddnerf_train_schedule: lr (nerf_helpers.py:211-245), Adam bias corrections and the annealed gaussian_smooth_factor
    (train_model.py:121-122,135-138) for a run of iterations, including the warm-up of the lr and the end of the anneal."""
    from ddnerf_b200 import _lib
    from ddnerf_b200.general_utils.nerf_helpers import learning_rate_decay
    lib = _lib.load()
    state = torch.zeros(2, device=DEV, dtype=torch.int64)
    hyper = torch.zeros(10, device=DEV)
    smooth0, final, fin = 1.7, 1.1, 30.0
    ds = (smooth0 - final) / fin
    sched = (ctypes.c_double * 13)(0.0005, 5e-6, 2000.0, 25.0, 0.01, 0.9, 0.999, 1e-8, 0.25, smooth0, ds, final, fin)
    for start in (0, 1990):
        state.copy_(torch.tensor([start, start + 3]))
        for k in range(45):
            _lib.check(lib.ddnerf_train_schedule(state.data_ptr(), hyper.data_ptr(), sched, None), "train_schedule")
            i, t = start + k, start + 3 + k + 1
            h = hyper.cpu().double()
            lr = learning_rate_decay(i, 0.0005, 5e-6, 2000, lr_delay_steps=25, lr_delay_mult=0.01)
            assert abs(h[0].item() - lr) <= 1e-6 * lr
            assert abs(h[4].item() - (1 - 0.9 ** t)) < 1e-6 and abs(h[5].item() - math.sqrt(1 - 0.999 ** t)) < 1e-6
            assert abs(h[6].item() - 0.25) < 1e-7
            want = smooth0 - ds * i if i < fin else final
            assert abs(h[7].item() - want) < 1e-6
            assert h[8].item() == float(torch.tensor(1 - 0.9, dtype=torch.float32)) and h[9].item() == float(torch.tensor(1 - 0.999, dtype=torch.float32))
        assert state.cpu().tolist() == [start + 45, start + 3 + 45]


@pytest.mark.parametrize("pname", ["config_blender_mipnerf", "config_360"])
def test_trainer_graph_queued_without_sync_matches_eager(pname):
    """24 graphed iterations queued back to back with NO host synchronisation (what bench.py's resident loop and any driver
    that does not read the loss every iteration do) walk the eager trajectory: every replay computes its own learning
    rate, bias corrections and smoothing factor on the device (round 1 read them from a pinned host buffer that the host
    had already overwritten for later iterations)."""
    from oracle import ddnerf_oracle as orc
    from ddnerf_b200.config import preset
    from ddnerf_b200.models import models as M
    from ddnerf_b200.rays import synth_rays
    from ddnerf_b200.trainer import Trainer
    dev = torch.device(DEV)
    N, s0, s1, n_iter = 256, 16, 16, 24
    ro, rd, rad, near, far = synth_rays("blender" if "blender" in pname else "360", N, seed=4)
    g = torch.Generator().manual_seed(1)
    target = torch.rand(N, 3, generator=g)
    rnd = dict(t_rand=torch.rand(N, s0 + 1, generator=g), noise0=torch.randn(N, s0, generator=g),
               u_rand=torch.rand(N, s1 + 1, generator=g), noise1=torch.randn(N, s1, generator=g))
    finals = []
    for use_graph in (False, True):
        cfg, _ = preset(pname, num_coarse=s0, num_fine=s1)
        cfg.train_params.finnish_smooth = 16          # the anneal ends inside the run
        is_dd = cfg.nerf.type == "DDNerfModel"
        model = getattr(M, cfg.nerf.type)(cfg)
        model.coarse.load_state_dict(orc.init_mlp_params(is_dd, seed=21))
        model.coarse.mlp_mode = "bf16"
        if is_dd:
            model.fine.load_state_dict(orc.init_mlp_params(False, seed=22))
            model.fine.mlp_mode = "bf16"
        model.to(dev)
        model.randoms = {k: v.to(dev) for k, v in rnd.items()}
        tr = Trainer(model, train_iters=100, use_graph=use_graph)      # short schedule: lr moves visibly every iteration
        args = [t.to(dev) for t in (ro, rd, rad, target)]
        losses = []
        for _ in range(n_iter):
            loss, mse = tr.step(*args)
            losses.append(loss.clone())                                # device-side copy, no sync
        torch.cuda.synchronize()
        assert use_graph == (tr._graph is not None)
        assert abs(cfg.train_params.gaussian_smooth_factor - cfg.train_params.final_smooth) < 1e-12
        finals.append((torch.stack(losses).cpu(), torch.cat([b.flat.clone() for b in tr.buckets]).cpu(), tr))
    (l0, w0, _), (l1, w1, tr) = finals
    assert (l0 - l1).abs().max().item() < 2e-5, (l0, l1)
    # Adam normalises every element: one whose gradient is round-off noise may move by a learning-rate step either way
    assert (w0 - w1).abs().max().item() < 2e-3
    assert torch.nn.functional.cosine_similarity((w0 - w0.mean()).double(), (w1 - w1.mean()).double(), dim=0).item() > 0.99999
    assert tr._sched_state.cpu().tolist() == [n_iter, n_iter]


def test_trainer_checkpoint_resume_continues_the_trajectory():
    """Trainer.state_dict() / resume(): 4 iterations, checkpoint, 4 more == 8 iterations in one go (weights, Adam moments,
    step counts, lr schedule position), through the graphed path."""
    from oracle import ddnerf_oracle as orc
    from ddnerf_b200.config import preset
    from ddnerf_b200.models import models as M
    from ddnerf_b200.rays import synth_rays
    from ddnerf_b200.trainer import Trainer
    dev = torch.device(DEV)
    N, s0, s1 = 256, 16, 16
    ro, rd, rad, near, far = synth_rays("360", N, seed=4)
    g = torch.Generator().manual_seed(1)
    target = torch.rand(N, 3, generator=g)
    rnd = dict(t_rand=torch.rand(N, s0 + 1, generator=g), noise0=torch.randn(N, s0, generator=g),
               u_rand=torch.rand(N, s1 + 1, generator=g), noise1=torch.randn(N, s1, generator=g))
    args = [t.to(dev) for t in (ro, rd, rad, target)]

    def fresh():
        cfg, _ = preset("config_360", num_coarse=s0, num_fine=s1)
        model = M.DDNerfModel(cfg)
        model.coarse.load_state_dict(orc.init_mlp_params(True, seed=21))
        model.fine.load_state_dict(orc.init_mlp_params(False, seed=22))
        for net in (model.coarse, model.fine):
            net.mlp_mode = "bf16"
        model.to(dev)
        model.randoms = {k: v.to(dev) for k, v in rnd.items()}
        return model

    a = Trainer(fresh(), train_iters=100, use_graph=True)
    for _ in range(8):
        a.step(*args)
    b = Trainer(fresh(), train_iters=100, use_graph=True)
    for _ in range(4):
        b.step(*args)
    ck = b.state_dict()
    assert ck["iter"] == 3 and set(ck) >= {"model_1_state_dict", "model_2_state_dict", "optimizer_1_state_dict",
                                           "optimizer_2_state_dict"}
    c = Trainer(fresh(), train_iters=100, use_graph=True)
    assert c.resume(ck) == 4
    for _ in range(4):
        c.step(*args)
    torch.cuda.synchronize()
    wa = torch.cat([x.flat for x in a.buckets]).cpu()
    wc = torch.cat([x.flat for x in c.buckets]).cpu()
    assert [x.step for x in c.buckets] == [8, 8] and c.iter == 8
    assert (wa - wc).abs().max().item() < 2e-3
    assert torch.nn.functional.cosine_similarity((wa - wa.mean()).double(), (wc - wc.mean()).double(), dim=0).item() > 0.99999


@pytest.mark.parametrize("pname", ["config_blender_mipnerf", "config_blender"])
def test_gradient_accumulation_over_ray_chunks(pname):
    """A batch above cfg.nerf.train.chunksize: backward per chunk into the flat gradient bucket (Trainer.accumulate_chunks)
    gives the gradient, the loss and the weight update of the reference's order -- every chunk forward (models.py:53,160),
    one backward over all of them (train_model.py:170) -- while only one chunk's saved activations are alive at a time.
    700 rays in chunks of 256 (ragged last chunk), DDNeRF (blender row filter, per-chunk dp-loss means) and mip-NeRF."""
    from oracle import ddnerf_oracle as orc
    from ddnerf_b200.config import preset
    from ddnerf_b200.models import models as M
    from ddnerf_b200.rays import synth_rays
    from ddnerf_b200.trainer import Trainer
    dev = torch.device(DEV)
    N, s0, s1 = 700, 16, 16
    ro, rd, rad, near, far = synth_rays("blender", N, seed=6)
    g = torch.Generator().manual_seed(3)
    target = torch.rand(N, 3, generator=g)
    rnd = dict(t_rand=torch.rand(N, s0 + 1, generator=g), noise0=torch.randn(N, s0, generator=g),
               u_rand=torch.rand(N, s1 + 1, generator=g), noise1=torch.randn(N, s1, generator=g))
    res = []
    for accumulate in (False, True):
        cfg, _ = preset(pname, num_coarse=s0, num_fine=s1)
        cfg.nerf.train.chunksize = 256
        is_dd = cfg.nerf.type == "DDNerfModel"
        model = getattr(M, cfg.nerf.type)(cfg)
        model.coarse.load_state_dict(orc.init_mlp_params(is_dd, seed=21))
        model.coarse.mlp_mode = "bf16"
        if is_dd:
            model.fine.load_state_dict(orc.init_mlp_params(False, seed=22))
            model.fine.mlp_mode = "bf16"
        model.to(dev)
        model.randoms = {k: v.to(dev) for k, v in rnd.items()}
        tr = Trainer(model, train_iters=100, use_graph=False)
        tr.accumulate_chunks = accumulate
        torch.cuda.reset_peak_memory_stats()
        loss, mse = tr.step(*[t.to(dev) for t in (ro, rd, rad, target)])
        torch.cuda.synchronize()
        res.append((loss.item(), mse.cpu(), torch.cat([b.grad.clone() for b in tr.buckets]).cpu(),
                    torch.cat([b.flat.clone() for b in tr.buckets]).cpu(), torch.cuda.max_memory_allocated()))
    (l0, m0, g0, w0, mem0), (l1, m1, g1, w1, mem1) = res
    assert abs(l0 - l1) < 1e-5 and (m0 - m1).abs().max().item() < 1e-6
    assert ((g0 - g1).abs().max() / g0.abs().max()).item() < 1e-4           # fp32 accumulation order only
    assert torch.nn.functional.cosine_similarity(g0.double(), g1.double(), dim=0).item() > 0.9999999
    assert (w0 - w1).abs().max().item() < 1e-4
    assert mem1 < mem0                                                      # one chunk's saves instead of three


def test_trainer_arena_holds_both_networks():
    """Trainer places the flat buckets of both DDNeRF networks in ONE arena (parameters, gradients, Adam moments): the
    buckets are 256-byte-aligned slices, the parameters alias them, a step zeroes / all-reduces / updates the arena as one
    buffer and the padding between the buckets stays zero; differing Adam step counts fall back to per-bucket updates."""
    from oracle import ddnerf_oracle as orc
    from ddnerf_b200.config import preset
    from ddnerf_b200.models import models as M
    from ddnerf_b200.rays import synth_rays
    from ddnerf_b200.trainer import Trainer
    dev = torch.device(DEV)
    N, s0, s1 = 128, 16, 16
    cfg, _ = preset("config_blender", num_coarse=s0, num_fine=s1)
    model = M.DDNerfModel(cfg)
    model.coarse.load_state_dict(orc.init_mlp_params(True, seed=31))
    model.fine.load_state_dict(orc.init_mlp_params(False, seed=32))
    model.coarse.mlp_mode = model.fine.mlp_mode = "bf16"
    model.to(dev)
    tr = Trainer(model, train_iters=100, use_graph=False)
    assert tr.arena is not None and len(tr.buckets) == 2
    flat, grad, m, v = tr.arena
    n0, n1 = tr.buckets[0].flat.numel(), tr.buckets[1].flat.numel()
    off1 = (n0 + 63) // 64 * 64
    assert flat.numel() == off1 + (n1 + 63) // 64 * 64
    for b, off in zip(tr.buckets, (0, off1)):
        for whole, part in zip(tr.arena, (b.flat, b.grad, b.exp_avg, b.exp_avg_sq)):
            assert part.data_ptr() == whole.data_ptr() + 4 * off and part.data_ptr() % 256 == 0
        b.check_alias()
    w_before = flat.clone()
    ro, rd, rad, _, _ = synth_rays("blender", N, seed=6)
    target = torch.rand(N, 3, generator=torch.Generator().manual_seed(2))
    args = [t.to(dev) for t in (ro, rd, rad, target)]
    for _ in range(3):
        tr.step(*args)
    torch.cuda.synchronize()
    assert [b.step for b in tr.buckets] == [3, 3]
    pad = torch.ones(flat.numel(), dtype=torch.bool, device=dev)
    pad[:n0] = False
    pad[off1:off1 + n1] = False
    for t in tr.arena:
        assert t[pad].abs().max().item() == 0.0                     # zero gradients, zero moments, zero parameters
    for b in tr.buckets:
        assert b.grad.abs().max().item() > 0 and (b.flat - w_before[b.flat.data_ptr() // 4 - flat.data_ptr() // 4:][:b.flat.numel()]).abs().max().item() > 0
    # a bucket restored from a checkpoint with another step count: the update falls back to one launch per bucket
    tr.buckets[1].step = 7
    tr.step(*args)
    torch.cuda.synchronize()
    assert [b.step for b in tr.buckets] == [4, 8]
    for t in tr.arena:
        assert t[pad].abs().max().item() == 0.0


@pytest.mark.parametrize("pname", ["config_blender", "config_blender_mipnerf"])
def test_trainer_fp32_mode_gathers_into_the_arena(pname):
    """fp32 parity mode has no gradient sink: after backward the per-parameter .grad tensors are gathered into the buckets'
    slices of the gradient arena (FlatBucket.gather_grads) and the one Adam launch updates every network.  The gathered
    gradient equals what autograd leaves on an identical model, and the parameters move."""
    from oracle import ddnerf_oracle as orc
    from ddnerf_b200.config import preset
    from ddnerf_b200.models import models as M
    from ddnerf_b200.rays import synth_rays
    from ddnerf_b200.trainer import Trainer
    dev = torch.device(DEV)
    N, s0, s1 = 96, 16, 16
    ro, rd, rad, _, _ = synth_rays("blender", N, seed=8)
    g = torch.Generator().manual_seed(3)
    target = torch.rand(N, 3, generator=g)
    rnd = dict(t_rand=torch.rand(N, s0 + 1, generator=g), noise0=torch.randn(N, s0, generator=g),
               u_rand=torch.rand(N, s1 + 1, generator=g), noise1=torch.randn(N, s1, generator=g))

    def fresh():
        cfg, _ = preset(pname, num_coarse=s0, num_fine=s1)
        is_dd = cfg.nerf.type == "DDNerfModel"
        model = getattr(M, cfg.nerf.type)(cfg)
        model.coarse.load_state_dict(orc.init_mlp_params(is_dd, seed=51))
        if is_dd:
            model.fine.load_state_dict(orc.init_mlp_params(False, seed=52))
        model.to(dev)
        model.randoms = {k: v.to(dev) for k, v in rnd.items()}
        return cfg, model, is_dd

    args = [t.to(dev) for t in (ro, rd, rad, target)]
    # autograd on a plain model
    cfg, model, is_dd = fresh()
    model.record_distributions = False
    out = model.run_iter(*args[:3], mode="train", rgb_target=args[3])
    tp = cfg.train_params
    loss = sum(tp.loss_coeficients[j] * torch.nn.functional.mse_loss(out[j]["rgb"], args[3]) for j in range(2))
    if is_dd:
        loss = loss + tp.dp_coeficient * out[1]["dp_loss"].mean()
    loss.backward()
    nets = [model.coarse] + ([model.fine] if is_dd else [])
    want = [torch.cat([p.grad.reshape(-1) for p in net.parameters()]) for net in nets]
    # the Trainer on an identical model
    cfg, model, is_dd = fresh()
    tr = Trainer(model, train_iters=100, use_graph=False)
    assert tr.arena is not None and not any(b.sink for b in tr.buckets)
    w_before = [b.flat.clone() for b in tr.buckets]
    loss_t, _ = tr.step(*args)
    torch.cuda.synchronize()
    assert abs(loss_t.item() - loss.item()) < 1e-5
    for b, w, w0 in zip(tr.buckets, want, w_before):
        assert not b.sink and b.step == 1
        sc = w.abs().max().clamp(min=1e-12)
        assert ((b.grad - w).abs().max() / sc).item() < 1e-4
        assert (b.flat - w0).abs().max().item() > 0
        assert all(p.grad is None for p in b.params)

"""One training step of the reference's loop (train_model.py:132-177) on the CUDA path, with the
data-parallel extension of SURVEY.md section 8e.

The reference keeps 26 (+24) parameter tensors and steps two ``torch.optim.Adam`` objects over
them tensor by tensor.  Here each network's parameters are views into a flat fp32 bucket, and the
buckets of all networks (DDNeRF: coarse and fine) are slices of ONE arena with matching arenas for
the gradients and the two Adam moments: a step zeroes the gradient arena once, sums it across ranks
with a single NCCL all-reduce (rays shard across ranks, weights are replicated) and applies it with
one fused Adam launch.  The per-parameter ``nn.Parameter`` objects (names, shapes, ``state_dict``)
and the per-network optimizer states stay what the reference's checkpoints expect.
"""
import torch
import torch.distributed as dist

from . import ops
from .general_utils.nerf_helpers import learning_rate_decay  # noqa: F401  (re-exported)


class FlatBucket:
    """All parameters of one network as views into one contiguous fp32 buffer + Adam state."""

    def __init__(self, module, storage=None):
        """``storage``: (flat, grad, exp_avg, exp_avg_sq) slices of a larger allocation (`Trainer` places the buckets of
        both DDNeRF networks in one arena so that a step has ONE memset, ONE all-reduce and ONE Adam launch); None: own
        buffers."""
        self.module = module
        self.params = [p for p in module.parameters()]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        if storage is None:
            self.flat = torch.empty(total, device=dev, dtype=torch.float32)
            self.grad = torch.zeros_like(self.flat)
            self.exp_avg = torch.zeros_like(self.flat)
            self.exp_avg_sq = torch.zeros_like(self.flat)
        else:
            self.flat, self.grad, self.exp_avg, self.exp_avg_sq = storage
            for t in storage:
                if t.numel() != total or t.device != dev or t.dtype != torch.float32:
                    raise RuntimeError("ddnerf_b200: FlatBucket storage must be fp32 slices of the module's parameter count on its device")
        off = 0
        for p in self.params:
            n = p.numel()
            self.flat[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + n].view(p.shape)
            off += n
        self.step = 0
        self.sink = False

    @staticmethod
    def numel_of(module):
        return sum(p.numel() for p in module.parameters())

    def install_sink(self):
        """bf16 mode: let the weight-gradient kernel accumulate directly into this bucket's flat gradient (views in the
        (weights, biases) order of the kernel's parameter table).  Replaces, per step, ~26 gradient clones / adds, the
        zero fills of the per-call gradient buffers and the concatenation of `gather_grads` by one memset."""
        if getattr(self.module, "mlp_mode", None) != "bf16" or not hasattr(self.module, "_param_pairs"):
            return False
        off, where = 0, {}
        for p in self.params:
            where[id(p)] = (off, p.numel(), p.shape)
            off += p.numel()
        def view(p):
            o, n, shp = where[id(p)]
            return self.grad[o:o + n].view(shp)
        pairs = self.module._param_pairs()
        object.__setattr__(self.module, "_grad_sink", ([view(w) for w, _ in pairs], [view(b) for _, b in pairs]))
        self.sink = True
        return True

    def remove_sink(self):
        if self.sink:
            object.__setattr__(self.module, "_grad_sink", None)
            self.sink = False

    def begin_step(self):
        """Before the forward of a step: with the sink installed the gradient bucket is an accumulator."""
        if self.sink:
            self.grad.zero_()

    def gather_grads(self):
        """Copy the per-parameter .grad tensors into the flat gradient bucket (one cat kernel).  With the sink
        installed the kernels have already written there."""
        if self.sink:
            return
        torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self.params],
                  out=self.grad)
        for p in self.params:
            p.grad = None

    def mark_dirty(self):
        """The optimizer kernels write through raw pointers (no autograd version bump): tell the bf16 path
        that its packed weight images are stale."""
        st = getattr(self.module, "_tc_state", None)
        if st is not None:
            st.dirty = True

    def check_alias(self):
        """`module.to()` / `.float()` after construction re-allocates the parameters; Adam would then update only the
        flat buffer and the model would silently stop learning."""
        lo = self.flat.data_ptr()
        hi = lo + self.flat.numel() * 4
        for p in self.params:
            if not (lo <= p.data_ptr() < hi):
                raise RuntimeError("ddnerf_b200: a parameter no longer aliases the trainer's flat bucket (module.to() / "
                                   "load of a different device after Trainer construction?); build the Trainer last")

    # ---- torch.optim.Adam interchange (train_model.py:110-118, 249-258) ---------------------------------
    def state_dict(self, lr=0.0005):
        """This bucket's optimizer state in the layout of ``torch.optim.Adam.state_dict()`` over
        ``module.parameters()``: what the reference stores as ``optimizer_{1,2}_state_dict``."""
        state, off = {}, 0
        for i, p in enumerate(self.params):
            n = p.numel()
            if self.step > 0:
                state[i] = {"step": torch.tensor(float(self.step)),
                            "exp_avg": self.exp_avg[off:off + n].view(p.shape).clone(),
                            "exp_avg_sq": self.exp_avg_sq[off:off + n].view(p.shape).clone()}
            off += n
        group = {"lr": lr, "betas": (0.9, 0.999), "eps": 1e-08, "weight_decay": 0, "amsgrad": False, "maximize": False,
                 "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        """Inverse of ``state_dict``; accepts what ``torch.optim.Adam.state_dict()`` wrote for the same module."""
        state, off, steps = sd["state"], 0, set()
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        for i, p in enumerate(self.params):
            n = p.numel()
            st = state.get(i, state.get(str(i)))
            if st is not None:
                self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                steps.add(int(float(st["step"])))
            off += n
        if len(steps) > 1:
            raise RuntimeError(f"ddnerf_b200: per-parameter Adam step counts differ ({sorted(steps)}); one flat bucket "
                               "steps all parameters together")
        self.step = steps.pop() if steps else 0

    def adam(self, lr, grad_scale=1.0):
        self.check_alias()
        self.step += 1
        ops.adam_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, lr, self.step, grad_scale=grad_scale)
        self.mark_dirty()

    def adam_dev(self, hyper):
        """Same update with its scalars in device memory (graph capture); the caller advances `step`."""
        ops.adam_step_dev(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, hyper)


def shard_rows(n_rows, rank, world):
    """Contiguous pixel-row range [lo, hi) of one rank when a frame of `n_rows` image rows is rendered
    by `world` ranks (rendering splits rows, no collective)."""
    lo = (n_rows * rank) // world
    hi = (n_rows * (rank + 1)) // world
    return lo, hi


def allreduce_gradients(buckets, world):
    """The one collective of a data-parallel step: sum every flat gradient bucket across ranks (NCCL
    on GPUs, gloo in the CPU tests).  The 1/world factor is applied by the optimizer kernel."""
    if world > 1:
        for b in buckets:
            dist.all_reduce(b.grad, op=dist.ReduceOp.SUM)


class Trainer:
    """Drives ``model.run_iter`` + loss + backward + (all-reduce) + Adam, one call per iteration.

    ``use_graph=True`` captures the whole iteration (19 kernel launches for mip-NeRF, 25 for DDNeRF, and many more host-side
    dispatches) into ONE CUDA graph after two eager iterations and replays it from then on; the values that
    change every iteration -- learning rate, Adam bias corrections, the annealed ``gaussian_smooth_factor`` --
    are computed ON THE DEVICE by the graph's first node from a device-resident iteration counter
    (``ddnerf_train_schedule``), so a replay reads nothing the host mutates and any number of replays may be
    queued without a host sync.  With several ranks the NCCL all-reduce is captured into the same graph.  The
    graph is re-captured when ``pdf_padding`` flips (train_model.py:140-142) or the batch shape changes.  The
    results are those of the eager path: the same kernels in the same order."""

    GRAPH_WARMUP = 2

    def __init__(self, model, train_iters=None, distributed=None, use_graph=False):
        self.model = model
        self.cfg = model.cfg
        self.is_dd = self.cfg.nerf.type == "DDNerfModel"
        # One arena {parameters, gradients, Adam moments} for all networks (train_model.py:86-98 keeps two optimizers over
        # ~50 tensors): every bucket is a 256-byte-aligned slice of it, so a step zeroes, all-reduces and updates ONE buffer.
        nets = [model.coarse] + ([model.fine] if self.is_dd else [])   # train_model.py:93-98 second optimizer
        sizes = [FlatBucket.numel_of(m) for m in nets]
        starts, total = [], 0
        for n in sizes:
            starts.append(total)
            total += (n + 63) // 64 * 64
        dev0 = next(model.coarse.parameters()).device
        if all(p.device == dev0 for m in nets for p in m.parameters()):
            self.arena = tuple(torch.zeros(total, device=dev0, dtype=torch.float32) for _ in range(4))
            self.buckets = [FlatBucket(m, tuple(a[o:o + n] for a in self.arena)) for m, o, n in zip(nets, starts, sizes)]
        else:
            self.arena = None
            self.buckets = [FlatBucket(m) for m in nets]
        if train_iters is None:
            try:
                train_iters = int(self.cfg.experiment.train_iters)
            except Exception:
                train_iters = 200001
        self.train_iters = train_iters
        self.iter = 0
        self.distributed = dist.is_available() and dist.is_initialized() if distributed is None else distributed
        self.world = dist.get_world_size() if self.distributed else 1
        tp = self.cfg.train_params
        self._smooth0 = tp.gaussian_smooth_factor
        self._dsmooth = (tp.gaussian_smooth_factor - tp.final_smooth) / tp.finnish_smooth
        model.record_distributions = False
        self.use_graph = bool(use_graph)
        # several ranks: capture the NCCL all-reduce into the step graph (one graph per step); False = two graphs around
        # an eagerly launched all-reduce (round-1 behaviour)
        self.capture_collective = True
        # batches above cfg.nerf.train.chunksize rays: backward per chunk (see _body); False = the reference's order
        # (all chunks forward, one backward), which needs the saved activations of the whole batch at once
        self.accumulate_chunks = True
        self._counter_iter = -1
        self._consts = {}
        self._eager_calls = 0
        self._graph = None
        self._graph_tail = None
        self._graph_key = None
        self._static = None
        dev = self.buckets[0].flat.device
        if self.use_graph:
            # [lr, beta1, beta2, eps, 1 - beta1^t, sqrt(1 - beta2^t), grad_scale, gaussian_smooth_factor, 1 - beta1, 1 - beta2]
            self._hyper_dev = torch.zeros(10, device=dev, dtype=torch.float32)
            self._sched_state = torch.zeros(2, device=dev, dtype=torch.int64)      # {iteration, Adam steps taken}

    # ---- checkpoints in the reference's format (train_model.py:248-263 save, :77-81,110-118 resume) ---------------
    def state_dict(self, loss=None, psnr=None):
        """The dict ``train_model.py`` saves every ``save_every`` iterations (after iteration ``self.iter - 1``)."""
        ck = {"iter": self.iter - 1, "model_1_state_dict": self.model.coarse.state_dict(),
              "optimizer_1_state_dict": self.buckets[0].state_dict(self.lr(max(self.iter - 1, 0))), "loss": loss, "psnr": psnr}
        if self.is_dd:
            ck["model_2_state_dict"] = self.model.fine.state_dict()
            ck["optimizer_2_state_dict"] = self.buckets[1].state_dict(self.lr(max(self.iter - 1, 0)))
        return ck

    def resume(self, checkpoint):
        """Continue from a checkpoint written by the reference's driver or by ``state_dict``: weights, Adam moments and
        step counts, ``start_iter = iter + 1`` and the ``pdf_padding`` fix-up of train_model.py:116-118."""
        self.model.load_weights_from_checkpoint(checkpoint)
        for k, b in enumerate(self.buckets):
            sd = checkpoint.get(f"optimizer_{k + 1}_state_dict")
            if sd is not None:
                b.load_state_dict(sd)
            b.mark_dirty()
        self.iter = int(checkpoint["iter"]) + 1
        if self.iter > self.cfg.train_params.max_pdf_pad_iters:
            self.cfg.train_params.pdf_padding = False
        self._graph = None                                     # re-capture: the device counters restart from here
        return self.iter

    def lr(self, i):
        return learning_rate_decay(i, 0.0005, 5e-6, self.train_iters, lr_delay_steps=2500, lr_delay_mult=0.01)

    def _schedule(self, i):
        """Per-iteration host scalars of the driver loop (train_model.py:135-150)."""
        tp = self.cfg.train_params
        smooth = self._smooth0 - self._dsmooth * i if i < tp.finnish_smooth else tp.final_smooth
        if i == tp.max_pdf_pad_iters:
            tp.pdf_padding = False
        return smooth, self.lr(i)

    def _body(self, ray_origins, ray_directions, ray_rad, target, lr, hyper=None, collective_inside=True):
        """One iteration on the current stream: run_iter, losses, backward, gradient all-reduce, Adam.

        Batches larger than ``cfg.nerf.train.chunksize`` rays are walked chunk by chunk with the backward of each chunk
        run right after its forward (gradient accumulation into the flat bucket): the reference forwards every chunk
        first (models.py:53,160) and back-propagates through all of them at once (train_model.py:170), which keeps the
        saved activations of ALL chunks alive -- 5.4 KB per sample row in bf16 mode, 90 GB for 64 K rays x 256 samples.
        The sums are the same: mse over the batch = sum_c (n_c / N) mse_c, dp_loss.mean() = (1 / n_chunks) sum_c dp_c."""
        tp = self.cfg.train_params
        self.model.train()
        for b in self.buckets:
            if not b.sink:
                b.install_sink()                                 # (no-op unless the network runs in bf16 mode)
        if self.arena is not None and all(b.sink for b in self.buckets):
            self.arena[1].zero_()                                # the gradient accumulators of all networks: one memset
        else:
            for b in self.buckets:
                b.begin_step()
        coef = tp.loss_coeficients
        chunk = int(self.cfg.nerf.train.chunksize)
        ro, rd, rad = ray_origins.reshape(-1, 3), ray_directions.reshape(-1, 3), ray_rad.reshape(-1, 1)
        tgt = target.reshape(-1, 3)
        N = ro.shape[0]
        spans = [(lo, min(lo + chunk, N)) for lo in range(0, N, chunk)] if (self.accumulate_chunks and N > chunk) else [(0, N)]
        loss = mse = None
        for lo, hi in spans:
            frac = (hi - lo) / N
            self.model._rand_base = lo                           # (injected random tensors of the parity tests: this chunk's rows)
            out = self.model.run_iter(ro[lo:hi], rd[lo:hi], rad[lo:hi], mode="train", rgb_target=tgt[lo:hi])
            mse3, g0, g1 = ops.mse_loss_and_grad(out[0]["rgb"], out[1]["rgb"], tgt[lo:hi], coef[0] * frac, coef[1] * frac)
            mse_c, loss_c = mse3[:2], mse3[2]                    # (views: the kernel also wrote the weighted sum)
            tensors, grads = [out[0]["rgb"], out[1]["rgb"]], [g0, g1]
            if self.is_dd:                                       # train_model.py:163-167
                dp = out[1]["dp_loss"]
                dp = dp.reshape(()) if dp.numel() == 1 else dp.mean()    # (.mean() of the [1] tensor: a view, no launch)
                w_dp = tp.dp_coeficient / len(spans)
                loss_c = torch.add(loss_c, dp.detach(), alpha=w_dp)
                tensors.append(dp)
                grads.append(self._const(w_dp, dp.device))
            torch.autograd.backward(tensors, grads)
            del out, tensors, grads                              # the chunk's saved activations go back to the allocator
            loss = loss_c if loss is None else loss + loss_c
            if len(spans) == 1:
                mse = mse_c
            else:
                mse = mse_c * frac if mse is None else mse + mse_c * frac
        self.model._rand_base = 0
        for b in self.buckets:
            b.gather_grads()
        if hyper is not None and self.distributed and self.world > 1 and not collective_inside:
            return loss, mse                                     # two-graph mode: the collective and Adam follow outside
        self._allreduce()
        self._optimize(lr, hyper)
        return loss, mse

    def _allreduce(self):
        """The one collective of a data-parallel step: the gradient arena summed over ranks."""
        world = self.world if self.distributed else 1
        if world > 1 and self.arena is not None:
            dist.all_reduce(self.arena[1], op=dist.ReduceOp.SUM)
        else:
            allreduce_gradients(self.buckets, world)

    def _const(self, value, device):
        """0-dim fp32 constant on the device, created once (a cotangent that is the same every iteration)."""
        key = (float(value), str(device))
        t = self._consts.get(key)
        if t is None:
            if device.type == "cuda" and torch.cuda.is_current_stream_capturing():
                return torch.full((), float(value), device=device, dtype=torch.float32)
            t = self._consts[key] = torch.full((), float(value), device=device, dtype=torch.float32)
        return t

    def _optimize(self, lr, hyper=None):
        if self.arena is not None and len({b.step for b in self.buckets}) == 1:
            # one Adam launch over the arena (the padding between buckets has zero gradients and stays zero)
            for b in self.buckets:
                b.check_alias()
            flat, grad, m, v = self.arena
            if hyper is None:
                for b in self.buckets:
                    b.step += 1
                ops.adam_step(flat, grad, m, v, lr, self.buckets[0].step, grad_scale=1.0 / self.world)
                for b in self.buckets:
                    b.mark_dirty()
            else:
                ops.adam_step_dev(flat, grad, m, v, hyper)
            return
        for b in self.buckets:
            if hyper is None:
                b.adam(lr, grad_scale=1.0 / self.world)
            else:
                b.adam_dev(hyper)

    def step(self, ray_origins, ray_directions, ray_rad, target):
        """train_model.py:135-177 for one iteration.  Returns (loss, mse[2]) as device tensors (no
        host sync)."""
        i, tp = self.iter, self.cfg.train_params
        smooth, lr = self._schedule(i)
        if not self.use_graph or self._eager_calls < self.GRAPH_WARMUP:
            tp.gaussian_smooth_factor = smooth
            loss, mse = self._body(ray_origins, ray_directions, ray_rad, target, lr)
            self.iter += 1
            self._eager_calls += 1
            return loss, mse

        # ---- graphed iteration ------------------------------------------------------------------
        key = (bool(tp.pdf_padding), tuple(ray_origins.shape), tuple(target.shape), ray_origins.device)
        if self._graph is None or key != self._graph_key or self._counter_iter != i:
            self._capture(key, ray_origins, ray_directions, ray_rad, target, lr)
        for dst, src in zip(self._static["in"], (ray_origins, ray_directions, ray_rad, target)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        if self._graph_tail is not None:                       # data parallel without graph-captured NCCL: eager all-reduce
            self._allreduce()
            self._graph_tail.replay()
        for b in self.buckets:
            b.step += 1
            b.mark_dirty()
        self.iter += 1
        self._counter_iter = self.iter                         # the device counter advanced inside the graph
        tp.gaussian_smooth_factor = smooth                     # host copy for validation renders between iterations
        return self._static["loss"], self._static["mse"]

    def _schedule_dev(self):
        """Graph node: {lr, Adam bias corrections, grad_scale, smooth} of the device-side iteration counter."""
        import ctypes
        from . import _lib
        tp = self.cfg.train_params
        f32 = lambda x: float(torch.tensor(x, dtype=torch.float32))      # the eager kernel takes grad_scale as fp32
        sched = (ctypes.c_double * 13)(0.0005, 5e-6, float(self.train_iters), 2500.0, 0.01, 0.9, 0.999, 1e-8,
                                       f32(1.0 / self.world), float(self._smooth0), float(self._dsmooth),
                                       float(tp.final_smooth), float(tp.finnish_smooth))
        _lib.check(_lib.load().ddnerf_train_schedule(ctypes.c_void_p(self._sched_state.data_ptr()),
                                                     ctypes.c_void_p(self._hyper_dev.data_ptr()), sched, ops._stream()),
                   "train_schedule")

    def _capture(self, key, ray_origins, ray_directions, ray_rad, target, lr):
        tp = self.cfg.train_params
        self._static = {"in": [torch.empty_like(x) for x in (ray_origins, ray_directions, ray_rad, target)]}
        for dst, src in zip(self._static["in"], (ray_origins, ray_directions, ray_rad, target)):
            dst.copy_(src)
        for b in self.buckets:
            for p in b.params:
                p.grad = None
            b.mark_dirty()                                     # the graph re-packs the bf16 weight images every replay
            b.check_alias()
        steps = {b.step for b in self.buckets}
        if len(steps) != 1:
            raise RuntimeError(f"ddnerf_b200: the buckets' Adam step counts differ ({sorted(steps)})")
        # the device-side counters start where the host stands (first capture, re-capture, resume)
        self._sched_state.copy_(torch.tensor([self.iter, steps.pop()], dtype=torch.int64))
        self._counter_iter = self.iter
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        smooth_dev = self._hyper_dev[7]                        # 0-dim view: `sigmas * gaussian_smooth_factor` reads device memory
        saved = tp.gaussian_smooth_factor
        if isinstance(saved, torch.Tensor):
            saved = float(saved)
        nccl_in_graph = self.distributed and self.world > 1 and self.capture_collective

        def record(g, inside):
            try:
                with torch.cuda.graph(g):
                    self._schedule_dev()
                    tp.gaussian_smooth_factor = smooth_dev if self.is_dd else saved
                    return self._body(*self._static["in"], lr, hyper=self._hyper_dev, collective_inside=inside)
            finally:
                tp.gaussian_smooth_factor = saved

        try:
            loss, mse = record(graph, nccl_in_graph)
        except Exception:
            if not nccl_in_graph:
                raise
            # this NCCL / driver combination refuses to capture the collective: two graphs around an eager all-reduce
            self.capture_collective = nccl_in_graph = False
            for b in self.buckets:
                for p in b.params:
                    p.grad = None
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            loss, mse = record(graph, False)
        self._static["loss"], self._static["mse"] = loss, mse
        self._graph_tail = None
        if self.distributed and self.world > 1 and not nccl_in_graph:   # the collective launched eagerly between two graphs
            self._allreduce()
            torch.cuda.synchronize()
            tail = torch.cuda.CUDAGraph()
            with torch.cuda.graph(tail, pool=graph.pool()):
                self._optimize(lr, hyper=self._hyper_dev)
            self._graph_tail = tail
        # capture ran no kernels: the counters still hold this iteration
        self._graph, self._graph_key = graph, key

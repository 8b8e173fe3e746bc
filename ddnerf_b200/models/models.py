"""Mirror of the reference's ``models/models.py``: ``GeneralMipNerfModel`` (mip-NeRF) and
``DDNerfModel`` with the reference's public surface (``run_iter``, ``predict``, ``run_network``,
``get_rays_batches``, ``to``, ``load_weights_from_checkpoint``, ``train``, ``eval``, ``.coarse``,
``.fine``, ``.cfg``) and output-dict contract, on the CUDA operators of this package.

Differences in mechanism (not in results): ``run_network`` makes one fused encode+MLP call per
ray chunk instead of cast_rays -> IPE -> dir-enc -> cat -> per-minibatch 13-layer loops; the
samplers / compositor / dp-loss are single kernels.  Random tensors the reference draws inside the
path can be injected through ``self.randoms`` (dict with any of ``t_rand``, ``noise0``,
``u_rand``, ``noise1``; consumed per chunk) -- used by the parity tests.
"""
import torch

from . import base_architectures
from .samplers import *  # noqa: F401,F403
from .samplers import sample_first_cycle, sample_pdf, sample_pdf_with_mu_sigma
from .dd_utils import estimate_dp_loss
from ..general_utils.volume_rendering_utils import volume_render_radiance_field, _is_blender
from .. import ops
from ..general_utils.nerf_helpers import get_embedding_function, get_minibatches
from ..general_utils.math_utils import approximate_cdf, integrated_pos_enc


def _is_blender_type(cfg):
    """dd_utils.py:12 (the row filter of the dp-loss looks at the dataset type only)."""
    return cfg.dataset.type.lower() == "blender"


class GeneralMipNerfModel(torch.nn.Module):
    """models.py:9-184."""

    def __init__(self, cfg, backbone="MipNeRFModel"):
        super().__init__()
        self.coarse = getattr(base_architectures, backbone)(
            hidden_size=cfg.nerf.coarse_hidden_size, max_ipe_deg=16, num_encoding_fn_dir=4, include_input_xyz=False,
            include_input_dir=True, use_viewdirs=True)
        self.fine = self.coarse                                   # models.py:28 shared weights
        self.encode_position_fn = integrated_pos_enc
        self.cfg = cfg
        self.encode_direction_fn = get_embedding_function(num_encoding_functions=4, include_input=True,
                                                          log_sampling=True)
        self.randoms = None
        # models.py:292-300 records boolean-masked (data-dependent length) tensors every pass, which
        # forces a device sync; a trainer that does not log them can switch this off.
        self.record_distributions = True

    # ---- orchestration -------------------------------------------------------------------
    def run_iter(self, ray_origins, ray_directions, ray_rad, mode="train", depth_analysis_validation=False,
                 rgb_target=None):
        """models.py:40-73."""
        shape_rgb = ray_directions.shape
        shape_depth = ray_directions.shape[:-1]
        batches = self.get_rays_batches(ray_origins, ray_directions, ray_rad, mode)
        if rgb_target is not None:
            rgb_targets = get_minibatches(rgb_target.view((-1, 3)), chunksize=getattr(self.cfg.nerf, mode).chunksize)
        else:
            rgb_targets = [None for _ in batches]
        pred = []
        offset = 0
        for batch, tgt in zip(batches, rgb_targets):
            self._chunk_rows = (offset, offset + batch.shape[0])
            pred.append(self.predict(batch, mode, depth_analysis_validation, tgt))
            offset += batch.shape[0]
        output = pred[0]
        for i in range(1, len(pred)):
            for j in range(len(output)):
                for key in pred[i][j].keys():
                    if (pred[i][j][key] is not None) and (pred[i][j][key] is not False):
                        output[j][key] = torch.cat((output[j][key], pred[i][j][key]), dim=0)
        if mode == "validation" and not depth_analysis_validation:
            for i in range(len(output)):
                output[i]["rgb"] = output[i]["rgb"].view(shape_rgb)
                for k in ("disp", "acc", "depth"):
                    output[i][k] = output[i][k].view(shape_depth)
                if output[i].get("corrected_disp_map") is not None:
                    output[i]["corrected_disp_map"] = output[i]["corrected_disp_map"].view(shape_depth)
        return output

    def _rnd(self, key):
        """Injected random tensor for the current ray chunk, or None (= draw like the reference)."""
        if not self.randoms or self.randoms.get(key) is None:
            return None
        lo, hi = getattr(self, "_chunk_rows", (0, None))
        base = getattr(self, "_rand_base", 0)              # row offset of this run_iter call inside the injected tensors
        return self.randoms[key][base + lo: None if hi is None else base + hi]

    def _mode_cfg(self, mode):
        return getattr(self.cfg.nerf, mode)

    def _draw_randoms(self, N, mode, device):
        """The four random tensors the reference draws inside one predict() (samplers.py:57 t_rand [N,S0+1], :102 / :165
        u_rand [N,S1+1], volume_rendering_utils.py:31 noise0 [N,S0] and noise1 [N,S1]): injected ones are used as they are,
        the rest comes from ONE torch.rand and ONE torch.randn call per chunk (contiguous halves of a flat draw) instead
        of four generator launches.  None where the mode does not use the tensor."""
        mcfg = self._mode_cfg(mode)
        S0, S1 = mcfg.num_coarse, mcfg.num_fine
        r = {k: self._rnd(k) for k in ("t_rand", "u_rand", "noise0", "noise1")}
        uni = [(k, n) for k, n in (("t_rand", S0 + 1), ("u_rand", S1 + 1)) if bool(mcfg.perturb) and r[k] is None]
        nor = [(k, n) for k, n in (("noise0", S0), ("noise1", S1)) if mcfg.radiance_field_noise_std > 0.0 and r[k] is None]
        for group, draw in ((uni, torch.rand), (nor, torch.randn)):
            if not group:
                continue
            flat = draw(N * sum(n for _, n in group), dtype=torch.float32, device=device)
            off = 0
            for k, n in group:
                r[k] = flat[off:off + N * n].view(N, n)
                off += N * n
        if not bool(mcfg.perturb):
            r["t_rand"] = r["u_rand"] = None
        if not mcfg.radiance_field_noise_std > 0.0:
            r["noise0"] = r["noise1"] = None
        return r

    def predict(self, ray_batch, mode, depth_analysis_validation, rgb_target=None):
        """models.py:75-114."""
        if depth_analysis_validation:
            raise NotImplementedError("depth_analysis_validation plots (math_utils.py:210-278) are out of scope; "
                                      "run them with the reference implementation")
        rd = ray_batch[..., 3:6]
        near, far = ray_batch[..., 7:8], ray_batch[..., 8:9]
        mcfg = self._mode_cfg(mode)
        ret = {}
        t_vals = weights = None
        rnd = self._draw_randoms(ray_batch.shape[0], mode, ray_batch.device)
        for i in range(2):
            if i == 0:
                t_vals = sample_first_cycle(self.cfg, near, far, mode, t_rand=rnd["t_rand"])
            else:
                t_vals = sample_pdf(t_vals, weights, mcfg.num_fine + 1, self.cfg, det=(mcfg.perturb == 0.0),
                                    rand=rnd["u_rand"]).detach()
            radiance_field = self.run_network(ray_batch, t_vals, self.coarse, mode)
            rgb, disp, acc, weights, depth, _, _ = volume_render_radiance_field(
                radiance_field, t_vals, rd, radiance_field_noise_std=mcfg.radiance_field_noise_std,
                white_background=mcfg.white_background, cfg=self.cfg, noise=rnd[f"noise{i}"], want_rgb=False)
            ret[i] = {"rgb": rgb, "disp": disp, "acc": acc, "weights": weights, "depth": depth}
            self._record_t(i, t_vals)
        return ret

    def _record_t(self, i, t_vals):
        """Debug hook of the parity tests: with ``self.keep_t_vals = True`` the fence-posts of both passes of every chunk
        are kept in ``self.last_t_vals[pass]`` (list over chunks); off by default."""
        if getattr(self, "keep_t_vals", False):
            if i == 0 and getattr(self, "_chunk_rows", (0, 0))[0] == 0:
                self.last_t_vals = {0: [], 1: []}
            self.last_t_vals[i].append(t_vals.detach())

    def run_network(self, ray_batch, t_vals, network, mode):
        """models.py:117-142: [N,12] rays + [N,S+1] fence-posts -> [N,S,C] raw radiance field."""
        return network.forward_rays(ray_batch, t_vals, self.cfg.nerf.ray_shape)

    def get_rays_batches(self, ray_origins, ray_directions, ray_rad, mode):
        """models.py:144-162."""
        rays = ops.pack_rays(ray_origins, ray_directions, ray_rad, self.cfg.dataset.near, self.cfg.dataset.far)
        return get_minibatches(rays, chunksize=getattr(self.cfg.nerf, mode).chunksize)

    # ---- nn.Module plumbing the drivers rely on (models.py:164-184) -----------------------
    def to(self, device):
        self.coarse.to(device)
        self.fine.to(device)

    def load_weights_from_checkpoint(self, checkpoint):
        self.coarse.load_state_dict(checkpoint["model_1_state_dict"])
        if self.cfg.nerf.type != "GeneralMipNerfModel":
            self.fine.load_state_dict(checkpoint["model_2_state_dict"])

    def train(self, mode=True):
        self.coarse.train(mode)
        self.fine.train(mode)

    def eval(self):
        self.coarse.eval()
        self.fine.eval()


class DDNerfModel(GeneralMipNerfModel):
    """models.py:187-322."""

    def __init__(self, cfg):
        GeneralMipNerfModel.__init__(self, cfg, backbone="DepthMipNeRFModel")
        try:
            hidden_size_fine = cfg.nerf.fine_hidden_size
        except Exception:
            print("no nidden size params for fine model, set 256")
            hidden_size_fine = 256
        self.fine = base_architectures.MipNeRFModel(
            hidden_size=hidden_size_fine, max_ipe_deg=16, num_encoding_fn_dir=4, include_input_xyz=False,
            include_input_dir=True, use_viewdirs=True)

    def predict(self, ray_batch, mode, depth_analysis_validation, rgb_target=None):
        if depth_analysis_validation:
            raise NotImplementedError("depth_analysis_validation plots (math_utils.py:210-278) are out of scope; "
                                      "run them with the reference implementation")
        rd = ray_batch[..., 3:6]
        near, far = ray_batch[..., 7:8], ray_batch[..., 8:9]
        mcfg = self._mode_cfg(mode)
        tp = self.cfg.train_params
        ret = {}

        # ---- pass 0: coarse network with the depth-distribution head (models.py:222-273).  Everything after the
        # network -- sigmoid of the (mu, sigma) head, regulariser sums, compositing with the corrected depth -- is ONE
        # kernel (ops.composite_dd); the tails Phi((0-mu)/sigma), Phi((1-mu)/sigma) and their smoothed versions are
        # evaluated per cell inside the resampler / dp-loss kernels that consume them.
        rnd = self._draw_randoms(ray_batch.shape[0], mode, ray_batch.device)
        t0 = sample_first_cycle(self.cfg, near, far, mode, t_rand=rnd["t_rand"])
        rf0 = self.run_network(ray_batch, t0, self.coarse, mode)
        std = mcfg.radiance_field_noise_std
        rgb, disp, acc, w0, depth, cdisp, mus, sigmas, regs = ops.composite_dd(
            rf0, t0, rd, rnd["noise0"], std, mcfg.white_background, _is_blender(self.cfg), tp.dist_reg_coeficient)
        mus_loss, sig_loss, mus_reg, sig_reg = regs[0:1], regs[1:2], regs[2:3], regs[3:4]
        if self.record_distributions:                           # models.py:292-300 (mask from pass 0)
            smoothed_sigmas = sigmas * tp.gaussian_smooth_factor
            sel = (w0 / torch.sum(w0, dim=-1, keepdim=True)) > 0.1
            rec = {"mus": mus[sel], "sigmas": sigmas[sel], "smoothed_sigmas": smoothed_sigmas[sel]}
        else:
            rec = {"mus": None, "sigmas": None, "smoothed_sigmas": None}
        ret[0] = {"rgb": rgb, "disp": disp, "acc": acc, "weights": w0, "depth": depth, **rec, "dp_loss": None,
                  "corrected_disp_map": cdisp, "mus_loss": mus_loss, "sig_loss": sig_loss, "mus_reg": mus_reg,
                  "sig_reg": sig_reg}

        # ---- pass 1: fine network on depth-distribution samples (models.py:225-237, 276-289)
        t1 = ops.sample_pdf_mu_sigma_fused(t0, w0, mus, sigmas, tp.gaussian_smooth_factor, mcfg.num_fine + 1,
                                           tp.pdf_padding, self.cfg.dataset.near, self.cfg.dataset.far, rnd["u_rand"])
        self._record_t(0, t0)
        self._record_t(1, t1)
        rf1 = self.run_network(ray_batch, t1, self.fine, mode)
        rgb1, disp1, acc1, w1, depth1, _, _ = volume_render_radiance_field(
            rf1, t1, rd, radiance_field_noise_std=mcfg.radiance_field_noise_std,
            white_background=mcfg.white_background, mus=None, cfg=self.cfg, noise=rnd["noise1"], want_rgb=False)
        # models.py:287-289: estimate_dp_loss(...) * (S1) + mus_reg + sig_reg, [1]; the finishing kernel of the loss adds the
        # terms and its backward kernel returns the cotangent of regs (no slice / mul / add launches either way)
        dp_loss = ops.dp_loss_total(t1.detach(), t0.detach(), w1.detach(), w0, mus, sigmas, None, None, _is_blender_type(self.cfg),
                                    regs, t1.shape[1] - 1)
        ret[1] = {"rgb": rgb1, "disp": disp1, "acc": acc1, "weights": w1, "depth": depth1, **rec, "dp_loss": dp_loss,
                  "corrected_disp_map": None}
        return ret

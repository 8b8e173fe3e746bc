"""Mirror of the reference's ``models/base_architectures.py``: same constructor arguments, same
parameter names and [out,in] shapes (checkpoints interchange), forward executed by the MLP
kernels instead of 13 ``nn.Linear`` calls."""
import os

import torch

from .. import ops


class _NerfMLP(torch.nn.Module):
    out_channels = 4

    def __init__(self, num_layers=8, hidden_size=256, skip_connect_every=4, max_ipe_deg=16, num_encoding_fn_dir=4,
                 include_input_xyz=False, include_input_dir=False, use_viewdirs=True):
        super().__init__()
        include_input_xyz = 3 if include_input_xyz else 0
        include_input_dir = 3 if include_input_dir else 0
        self.dim_xyz = include_input_xyz + 2 * 3 * max_ipe_deg
        self.dim_dir = include_input_dir + 2 * 3 * num_encoding_fn_dir
        self.number_of_viewdir_layers = 1
        self.use_viewdirs = use_viewdirs
        if (hidden_size, self.dim_xyz, self.dim_dir, bool(use_viewdirs)) != (256, 96, 27, True):
            raise NotImplementedError(
                "ddnerf_b200 implements the architecture every shipped config uses: hidden 256, 96 IPE + 27 "
                f"direction features, view directions on (got hidden={hidden_size}, dim_xyz={self.dim_xyz}, "
                f"dim_dir={self.dim_dir}, use_viewdirs={use_viewdirs})")
        # construction order = the reference's (base_architectures.py:22-37), so seeded init matches
        self.layers_xyz = torch.nn.ModuleList()
        self.layers_xyz.append(torch.nn.Linear(self.dim_xyz, hidden_size))
        for i in range(1, 8):
            self.layers_xyz.append(torch.nn.Linear(self.dim_xyz + hidden_size if i == 5 else hidden_size, hidden_size))
        self.fc_feat = torch.nn.Linear(hidden_size, hidden_size)
        self.fc_alpha = torch.nn.Linear(hidden_size, 1)
        self.layers_dir = torch.nn.ModuleList()
        self.layers_dir.append(torch.nn.Linear(hidden_size + self.dim_dir, 128))
        self.fc_rgb = torch.nn.Linear(128, 3)
        self.relu = torch.nn.functional.relu
        # "fp32": CUDA-core SGEMM chain (1e-3 parity mode); "bf16": fused tcgen05 kernel
        self.mlp_mode = os.environ.get("DDNERF_MLP_MODE", "fp32")

    def _param_pairs(self):
        mods = list(self.layers_xyz) + [self.fc_feat, self.fc_alpha, self.layers_dir[0], self.fc_rgb]
        if self.out_channels == 6:
            mods.append(self.fc_mu_sigma)
        return [(m.weight, m.bias) for m in mods]

    def forward(self, x):
        """x [rows,123] -> [rows,4|6] (base_architectures.py:40-61 / 103-126)."""
        lead = x.shape[:-1]
        out = ops.mlp_f32(self._param_pairs(), self.out_channels, x=x.reshape(-1, x.shape[-1]))
        return out.reshape(*lead, self.out_channels)

    def forward_rays(self, rays, t_vals, ray_shape="cone"):
        """Fused encode + MLP: rays [N,12], t_vals [N,S+1] -> [N,S,4|6].  What run_network calls."""
        N, S = rays.shape[0], t_vals.shape[1] - 1
        if self.mlp_mode == "bf16":
            from .. import mlp_tc
            out = mlp_tc.mlp_bf16(self, rays, t_vals, ray_shape)
        else:
            out = ops.mlp_f32(self._param_pairs(), self.out_channels, rays=rays, t_vals=t_vals, ray_shape=ray_shape)
        return out.reshape(N, S, self.out_channels)


class MipNeRFModel(_NerfMLP):
    """base_architectures.py:3-61 -> [r,g,b,density]."""
    out_channels = 4


class DepthMipNeRFModel(_NerfMLP):
    """base_architectures.py:64-126 -> [r,g,b,density,raw_mu,raw_sigma]."""
    out_channels = 6

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.fc_mu_sigma = torch.nn.Linear(128, 2)

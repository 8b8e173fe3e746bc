"""Config records for the hot path.

``make_cfg`` builds a CfgNode with the layout of the reference's YAML files
(configs/config_blender.yml) restricted to the keys the path reads, with the
shipped defaults; ``preset`` returns the six shipped configurations plus the BASELINE.json
workload overrides.
"""
from .general_utils.cfgnode import CfgNode


def make_cfg(model="DDNerfModel", dataset_type="blender", basedir="synthetic", near=2.0, far=6.0, num_coarse=32,
             num_fine=32, chunksize=16384, num_random_rays=2048, noise_std=1.0, white_background=False, lindisp=False,
             pdf_padding=True, gaussian_smooth_factor=1.7, final_smooth=1.1, dist_reg_coeficient=0.02,
             loss_coeficients=(1, 1), dp_coeficient=0.1, ray_shape="cone", ndc_rays=False, lr=1.0e-3, seed=42):
    def mode(perturb):
        return {"chunksize": chunksize, "perturb": perturb, "num_coarse": num_coarse, "num_fine": num_fine,
                "white_background": white_background, "radiance_field_noise_std": noise_std, "lindisp": lindisp}
    train = mode(True)
    train["num_random_rays"] = num_random_rays
    return CfgNode({
        "experiment": {"id": "synthetic", "randomseed": seed},
        "train_params": {"pdf_padding": pdf_padding, "max_pdf_pad_iters": 20000,
                         "gaussian_smooth_factor": gaussian_smooth_factor, "final_smooth": final_smooth,
                         "finnish_smooth": 150000, "depth_analysis_rays": False,
                         "dist_reg_coeficient": dist_reg_coeficient, "set_automatic_dist_reg_coeficient": True,
                         "loss_coeficients": list(loss_coeficients), "dp_coeficient": dp_coeficient},
        "dataset": {"type": dataset_type, "basedir": basedir, "ndc_rays": ndc_rays, "near": near, "far": far,
                    "combined_sampling_method": False, "combined_split": 2, "normalize_poses": False,
                    "normalize_factor": 5},
        "optimizer": {"type": "Adam", "lr": lr},
        "nerf": {"type": model, "coarse_hidden_size": 256, "fine_hidden_size": 256, "ray_shape": ray_shape,
                 "train": train, "validation": mode(False)},
    })


def auto_dist_reg(num_coarse):
    """train_model.py:124-125."""
    return min(max(1 / num_coarse, 0.01), 0.12)


# name -> (make_cfg kwargs, ray preset of ddnerf_b200.rays.frame)
PRESETS = {
    # the six shipped YAML files
    "config_blender": (dict(model="DDNerfModel", dataset_type="blender"), "blender"),
    "config_blender_mipnerf": (dict(model="GeneralMipNerfModel", dataset_type="blender", loss_coeficients=(1, 0.1)), "blender"),
    "config_ff": (dict(model="DDNerfModel", dataset_type="LLFF", near=0.0, far=1.0, num_coarse=16, num_fine=16,
                       dist_reg_coeficient=0.1, ndc_rays=True), "ff"),
    "config_ff_mipnerf": (dict(model="GeneralMipNerfModel", dataset_type="LLFF", near=0.0, far=1.0, num_coarse=16,
                               num_fine=16, loss_coeficients=(1, 0.1), ndc_rays=True), "ff"),
    "config_360": (dict(model="DDNerfModel", dataset_type="REAL360", near=1.0 / 5, far=14.0 / 5), "360"),
    "config_360_mipnerf": (dict(model="GeneralMipNerfModel", dataset_type="REAL360", near=1.0 / 5, far=14.0 / 5,
                                loss_coeficients=(1, 0.1)), "360"),
}


def preset(name, **overrides):
    kwargs, ray_kind = PRESETS[name]
    kwargs = {**kwargs, **overrides}
    cfg = make_cfg(**kwargs)
    if cfg.train_params.set_automatic_dist_reg_coeficient and "dist_reg_coeficient" not in overrides:
        cfg.train_params.dist_reg_coeficient = auto_dist_reg(cfg.nerf.train.num_coarse)
    return cfg, ray_kind

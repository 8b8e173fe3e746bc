"""CPU suite: the C-ABI library loads, exports every symbol include/ddnerf_b200.h declares, and its
host-only logic (static kernel programs, work split of the weight-gradient kernel, size queries)
is consistent.  No device work."""
import ctypes
import os
import re

import numpy as np
import pytest

from ddnerf_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(REPO, "include", "ddnerf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ddnerf_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ddnerf_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in ddnerf_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.ddnerf_version() == 1
    # a refused call (null pointers) reports through the error string, without touching a device
    assert lib.ddnerf_mlp_tc_dw_plan(1024, 148, None, 0) != 0
    assert b"mlp_tc_dw_plan" in lib.ddnerf_last_error()


def test_kernel_programs_are_consistent():
    lib = _lib.load()
    _lib.check(lib.ddnerf_mlp_tc_program_check(), "program_check")
    assert lib.ddnerf_mlp_tc_wimg_bytes() % 16 == 0 and lib.ddnerf_mlp_tc_wimg_bytes() > 2 * 600_000
    assert lib.ddnerf_mlp_tc_bias_floats() == 16 * 256


@pytest.mark.parametrize("rows", [0, 1, 255, 256, 257, 4096 * 128, 65536 * 256])
def test_size_queries(rows):
    lib = _lib.load()
    items = lib.ddnerf_mlp_tc_items(rows)
    assert items == (rows + 255) // 256
    assert lib.ddnerf_mlp_tc_enc_bytes(rows) == items * 65536
    assert lib.ddnerf_mlp_tc_act_save_bytes(rows) == 10 * items * 2 * 65536
    assert lib.ddnerf_mlp_tc_mask_save_bytes(rows) == 10 * items * 2 * 128 * 32
    assert lib.ddnerf_mlp_f32_workspace_bytes(rows) >= 0


@pytest.mark.parametrize("rows,sms", [(256, 148), (592, 148), (19200, 148), (524288, 148), (524288, 132), (4096, 13)])
def test_dw_work_split_covers_every_tile_once(rows, sms):
    """backward_dw: every (layer-op, tile) pair belongs to exactly one work item; the tile line is cut into at most
    `sms` equal-cost pieces, and a piece crossing an op boundary is split there (at most 12 extra items)."""
    lib = _lib.load()
    buf = (ctypes.c_uint32 * (3 * 480))()
    n = lib.ddnerf_mlp_tc_dw_plan(rows, sms, buf, 480)
    assert 0 < n <= sms + 13
    tri = np.frombuffer(buf, dtype=np.uint32)[:3 * n].reshape(n, 3)
    n_tiles = 2 * ((rows + 255) // 256)
    cover = np.zeros((13, n_tiles), dtype=np.int32)
    for op, t0, t1 in tri:
        assert t1 > t0
        cover[op, t0:t1] += 1
    assert (cover == 1).all()
    if rows >= 19200:                                   # large problems: the pieces are equal to within one tile
        weight = np.array([88, 128, 128, 128, 128, 128, 88, 128, 128, 128, 112, 40, 48])
        total = weight.sum() * n_tiles
        assert total / sms > 0 and n <= sms + 12


def test_install_as_reference_aliases_the_driver_imports():
    """The reference's drivers do `from models import models` and `from general_utils import CfgNode`
    (train_model.py:4,14; eval_nerf.py:6,10): after install_as_reference() those names are this package."""
    import subprocess
    import sys
    code = (
        "import ddnerf_b200; ddnerf_b200.install_as_reference()\n"
        "from models import models\n"
        "from models.samplers import sample_pdf_with_mu_sigma\n"
        "from general_utils import CfgNode, mse2psnr, volume_render_radiance_field\n"
        "from general_utils.nerf_helpers import positional_encoding, get_minibatches\n"
        "assert models.DDNerfModel.__module__ == 'ddnerf_b200.models.models'\n"
        "assert volume_render_radiance_field.__module__ == 'ddnerf_b200.general_utils.volume_rendering_utils'\n"
        "from ddnerf_b200.config import preset\n"
        "cfg, _ = preset('config_blender')\n"
        "m = getattr(models, cfg.nerf.type)(cfg)\n"
        "names = [k for k, _ in m.coarse.named_parameters()]\n"
        "assert names[0] == 'layers_xyz.0.weight' and 'fc_mu_sigma.bias' in names and len(names) == 26\n"
        "print('ok')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=REPO, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]

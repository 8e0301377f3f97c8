"""The CPU oracle (oracle/vnlb_oracle.py + .c) against the golden fixtures made
by the reference's own code (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

import inputs as gin
from oracle import vnlb_oracle as orc


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _args(step, sigma=20.):
    return orc.get_args(orc.default_params(sigma), 3, step)


def test_color_bit_exact(golden_dir):
    g = _load(golden_dir, "color.npz")
    rgb = gin.color_inputs()
    yuv = orc.rgb2yuv(rgb)
    np.testing.assert_array_equal(yuv, g["yuv"])
    np.testing.assert_array_equal(orc.yuv2rgb(yuv), g["back"])


@pytest.mark.parametrize("shape", [(3, 3, 64, 64), (4, 3, 64, 64), (5, 3, 33, 41), (2, 3, 7, 9)])
def test_mask_init_bit_exact(golden_dir, shape):
    g = _load(golden_dir, "mask_init.npz")
    key = "x".join(map(str, shape))
    mask, ng = orc.init_mask(shape, _args(0))
    t, c, h, w = shape
    ref = np.unpackbits(g["mask_" + key])[: t * h * w].reshape(t, h, w)
    np.testing.assert_array_equal(mask, ref)
    assert ng == int(g["ngroups_" + key])


def test_mask_init_counts_survey():
    # 3x64x64 -> 824 set, 4x64x64 -> 1258
    assert orc.init_mask((3, 3, 64, 64), _args(0))[1] == 824
    assert orc.init_mask((4, 3, 64, 64), _args(0))[1] == 1258


def test_mask_update_bit_exact(golden_dir):
    g = _load(golden_dir, "mask_update.npz")
    inds, (T, C, H, W) = gin.mask_update_inputs()
    mask = np.ones((T, H, W), np.int8)
    orc.update_mask_inds(mask, inds, C)
    np.testing.assert_array_equal(mask, g["mask_after"])


def test_flat_areas(golden_dir):
    g = _load(golden_dir, "flat.npz")
    flat = orc.exec_flat_areas(gin.flat_inputs(), 0.2, 400.)
    np.testing.assert_array_equal(flat, g["flat"])
    assert flat.any() and not flat.all()


@pytest.mark.parametrize("step", [0, 1])
def test_bayes(golden_dir, step):
    g = _load(golden_dir, "bayes_step%d.npz" % (step + 1))
    pn, pb, flat = gin.bayes_inputs(step)
    out_n, out_b, rank_var = orc.bayes_denoise(pn, pb, flat, _args(step))
    # filtered patches within 1e-4 relative (north star), norm-wise per group
    for b in range(pn.shape[0]):
        err = np.linalg.norm(out_n[b] - g["noisy"][b]) / np.linalg.norm(g["noisy"][b])
        assert err < 1e-4, (b, err)
    np.testing.assert_allclose(out_n, g["noisy"], rtol=0, atol=2e-2)
    np.testing.assert_allclose(out_b, g["basic"], rtol=0, atol=1e-3)
    np.testing.assert_allclose(rank_var, g["rank_var"], rtol=1e-4)


def test_aggregation_bit_exact(golden_dir):
    g = _load(golden_dir, "agg.npz")
    p, inds, (T, C, H, W) = gin.agg_inputs()
    deno = np.zeros((T, C, H, W), np.float32)
    weights = np.zeros((T, H, W), np.float32)
    orc.agg_patches(deno, weights, p, inds)
    np.testing.assert_array_equal(deno, g["deno"])
    np.testing.assert_array_equal(weights, g["weights"])


def test_e2e_denoise(golden_dir):
    """Whole vnlb.denoise: oracle vs the reference's own run (default_params),
    tolerance of the north star: max-abs 1e-2, PSNR within 0.02 dB."""
    g = _load(golden_dir, "e2e.npz")
    e = gin.E2E
    clean = orc.synth_video(e["T"], e["H"], e["W"], e["seed"])
    noisy = orc.add_noise(clean, e["sigma"], e["seed"])
    torch.manual_seed(e["torch_seed"])
    deno, basic, _ = orc.denoise(noisy, e["sigma"])
    assert np.abs(basic - g["basic"]).max() < 1e-2
    assert np.abs(deno - g["deno"]).max() < 1e-2
    ps = [orc.compute_psnrs(a, clean).mean() for a in (noisy, basic, deno)]
    np.testing.assert_allclose(ps, g["psnrs"], atol=0.02)
    assert ps[2] > ps[1] > ps[0] + 8

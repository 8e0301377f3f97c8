"""Generates the golden fixtures under tests/golden/ by running the REFERENCE'S
OWN CODE (imported from /root/reference/lib, CPU, through the shims in
tests/golden/ref_shims/) on seeded inputs (tests/golden/inputs.py).

Run in the build container only (`python tests/golden/make_golden.py`);
/root/reference does not exist on the GPU box, so tests read the committed
.npz files, never the reference.

Shim recipe:
  * stub packages easydict / vpss / skimage / matplotlib on sys.path
    (vpss -> the CPU oracle search: the only stage whose reference code is
    absent, see oracle/vnlb_oracle.c);
  * torch.cuda.default_stream / synchronize / empty_cache -> no-ops
    (lib/vnlb/search/search.py:29-30,64; lib/vnlb/proc_nl.py:90,141);
  * vnlb.search_mask.mask.agg_boost (numba.cuda only, mask.py:104-187) -> a
    torch restatement of the same five-delta expansion;
  * vnlb.agg.comp_agg.compute_agg_batch -> direct call of exec_agg_simple
    (skips the GPU-only as_cuda_array wrappers, comp_agg.py:65-71);
  * get_params ("iphone" overrides, needs vpss's "needle" search) ->
    default_params (classic VNLB settings, lib/vnlb/params.py:11-50).
"""
import os
import sys
from types import SimpleNamespace

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/lib"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "ref_shims"))
sys.path.insert(0, REF)

import numpy as np
import torch

sys.path.insert(0, HERE)
import inputs as gin  # noqa: E402
from oracle import vnlb_oracle as orc  # noqa: E402


def import_reference():
    torch.cuda.default_stream = lambda *a, **k: SimpleNamespace(cuda_stream=0)
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.empty_cache = lambda *a, **k: None
    import vnlb
    import vnlb.search_mask.mask as rmask
    import vnlb.agg.comp_agg as ragg
    import vnlb.impl as rimpl
    import vnlb.params as rparams

    def agg_boost_torch(inds, t, c, h, w, cs_ptr):
        deltas = torch.tensor([[0, 0, 0], [0, 0, -1], [0, 0, 1], [0, 1, 0], [0, -1, 0]], dtype=torch.int64)
        lim = torch.tensor([t, h, w], dtype=torch.int64)
        valid_ind = torch.all((inds >= 0) & (inds < lim), 1)
        agg = inds[:, None, :] + deltas[None]
        ok = torch.all((agg >= 0) & (agg < lim), 2) & valid_ind[:, None]
        return agg[ok]

    rmask.agg_boost = agg_boost_torch

    def compute_agg_batch(deno, patches, inds, weights, vals, ivals, ps, ps_t, cs_ptr):
        ragg.exec_agg_simple(deno, patches, inds, weights, vals, ivals, ps, ps_t)

    ragg.compute_agg_batch = compute_agg_batch
    rimpl.get_params = lambda sigma, verbose=False: rparams.default_params(sigma, False)
    return vnlb


def main():
    vnlb = import_reference()
    from easydict import EasyDict as edict
    import vnlb.params as rparams
    import vnlb.deno as rdeno
    import vnlb.utils.color as rcolor
    import vnlb.utils.flat_areas as rflat
    import vnlb.search_mask.mask as rmask
    import vnlb.agg.comp_agg as ragg
    out = {}

    # ---- colour (A7, N2) ----
    rgb = gin.color_inputs()
    yuv = rcolor.rgb2yuv_cpp(torch.from_numpy(rgb.copy()))
    back = yuv.clone()
    rcolor.yuv2rgb_cpp(back)
    np.savez_compressed(os.path.join(HERE, "color.npz"), yuv=yuv.numpy(), back=back.numpy())

    # ---- mask init (A5) ----
    params = rparams.default_params(20.)
    masks = {}
    for shape in [(3, 3, 64, 64), (4, 3, 64, 64), (5, 3, 33, 41), (2, 3, 7, 9)]:
        args = rparams.get_args(params, 3, 0, "cpu")
        m, ng = rmask.init_mask(shape, args)
        key = "x".join(map(str, shape))
        masks["mask_" + key] = np.packbits(m.numpy().astype(bool))
        masks["ngroups_" + key] = np.int64(ng)
    np.savez_compressed(os.path.join(HERE, "mask_init.npz"), **masks)

    # ---- mask update / paste trick + boost (A10) ----
    inds, (T, C, H, W) = gin.mask_update_inputs()
    m = torch.ones((T, H, W), dtype=torch.int8)
    rmask.update_mask_inds(m, torch.from_numpy(inds), C, cs_ptr=0)
    np.savez_compressed(os.path.join(HERE, "mask_update.npz"), mask_after=m.numpy())

    # ---- flat areas (A11) ----
    x = gin.flat_inputs()
    patches = edict()
    patches.noisy = torch.from_numpy(x.copy())
    patches.flat = torch.zeros(x.shape[0], dtype=torch.bool)
    args = rparams.get_args(params, 3, 1, "cpu")
    rflat.update_flat_patch(patches, args)
    np.savez_compressed(os.path.join(HERE, "flat.npz"), flat=patches.flat.numpy())

    # ---- Bayes (B1-B7), with the intermediates of compute_cov_mat / denoise_eigvals / bayes_filter_coeff (B4-B6) ----
    import vnlb.deno.bayes_est as rbayes
    captured = {}
    orig_cov = rbayes.compute_cov_mat

    def capture_cov(pinput, rank):
        covMat, eigVals, eigVecs = orig_cov(pinput, rank)
        captured["cov"], captured["evals"] = covMat.clone(), eigVals.clone()
        captured["live"] = eigVals                      # mutated in place into the filter coefficients (bayes_est.py:41-42)
        return covMat, eigVals, eigVecs

    rbayes.compute_cov_mat = capture_cov
    for step in (0, 1):
        pn, pb, flat = gin.bayes_inputs(step)
        args = rparams.get_args(params, 3, step, "cpu")
        patches = edict()
        patches.noisy = torch.from_numpy(pn.copy())
        patches.basic = torch.from_numpy(pb.copy())
        patches.clean = None
        patches.flat = torch.from_numpy(flat.copy())
        patches.images = ["noisy", "basic", "clean"]
        patches.tensors = ["noisy", "basic", "clean", "flat"]
        rank_var = rdeno.denoise(patches, args, "bayes")
        np.savez_compressed(os.path.join(HERE, "bayes_step%d.npz" % (step + 1)),
                            noisy=patches.noisy.numpy(), basic=patches.basic.numpy(),
                            rank_var=rank_var.numpy())
        np.savez_compressed(os.path.join(HERE, "bayes_parts_step%d.npz" % (step + 1)),
                            cov=captured["cov"].numpy().astype(np.float32), evals=captured["evals"].numpy(),
                            coeff=captured["live"].numpy())
    rbayes.compute_cov_mat = orig_cov

    # ---- aggregation (G1) ----
    p, inds, (T, C, H, W) = gin.agg_inputs()
    deno = torch.zeros((T, C, H, W))
    weights = torch.zeros((T, H, W))
    patches = edict()
    patches.noisy = torch.from_numpy(p.copy())
    images = edict()
    images.deno, images.weights, images.vals = deno, weights, torch.zeros((T, H, W))
    bufs = edict()
    bufs.inds = torch.from_numpy(inds.copy())
    bufs.vals = torch.zeros(inds.shape)
    args = rparams.get_args(params, 3, 0, "cpu")
    ragg.agg_patches(patches, images, bufs, args, cs_ptr=0)
    np.savez_compressed(os.path.join(HERE, "agg.npz"), deno=images.deno.numpy(), weights=images.weights.numpy())

    # ---- end to end: reference vnlb.denoise (A1) with default_params: sigma 20 (e2e.npz), the noise levels of BASELINE
    #      configs[4] (sigma 10 / 50) and BASELINE configs[0] (3 x 64 x 64, sigma 20, the reference's CPU-runnable case) ----
    for name, e in gin.E2E_CASES.items():
        clean = orc.synth_video(e["T"], e["H"], e["W"], e["seed"])
        noisy = orc.add_noise(clean, e["sigma"], e["seed"])
        torch.manual_seed(e["torch_seed"])
        deno, basic, dt = vnlb.denoise(noisy.copy(), e["sigma"], gpuid=-1, verbose=False)
        deno, basic = deno.numpy(), basic.numpy()
        ps = [orc.compute_psnrs(a, clean).mean() for a in (noisy, basic, deno)]
        print("%s reference: %.2fs  psnr noisy %.3f basic %.3f deno %.3f" % (name, dt, *ps))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), deno=deno, basic=basic, psnrs=np.array(ps),
                            seconds=np.float64(dt))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()

"""Microbenchmarks of BASELINE.json configs[2] and configs[3] on the SHIPPED kernels (CUDA events, L2 flushed between
iterations).  Importable (bench.py puts the numbers in its JSON line) and runnable:
    python tools/microbench.py [nq] [ngroups]
configs[2]: sim_search, 7x7x2 patches, 27x27 window, +-4 frames, k=100 at 960x540 (16 frames), lattice pixels as queries
configs[3]: Bayes filter, groups of k=100 (step 1) / k=60 (step 2) patches of dim 7*7*2*3: cov + eig + Wiener (stack
            operator vnlb_bayes_filter: reads and writes the patch stacks, SURVEY 8d "materialised boundary")"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def _timeit(fn, flush, iters=5):
    ts = []
    for _ in range(iters + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[2:]))


def search_microbench(dev="cuda:0", nq=65536, fp32_peak_tflops=74.4, flush=None):
    import vnlb_b200
    from vnlb_b200 import color, search, synth
    from vnlb_b200 import mask as gmask
    flush = flush if flush is not None else torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    T, H, W = 16, 540, 960
    clean = synth.synth_video(T, H, W)
    yuv = color.rgb2yuv(torch.from_numpy(synth.add_noise(clean, 20.)).to(dev))
    params = vnlb_b200.get_params(20.)
    params["sizeSearchTimeFwd"] = [4, 4]
    params["sizeSearchTimeBwd"] = [4, 4]
    params["nSimilarPatches"] = [100, 100]
    out = {}
    for step in (0, 1):
        a = vnlb_b200.get_args(params, 3, step, dev)
        m, nset = gmask.init_mask(yuv.shape, a, dev)
        q = torch.nonzero(m)
        q = q[:: max(1, q.shape[0] // nq)][:nq].contiguous()
        k = a.npatches
        vals = torch.empty((q.shape[0], k), device=dev)
        inds = torch.empty((q.shape[0], k), dtype=torch.int64, device=dev)
        ms = _timeit(lambda: search.exec_sim_search_burst(yuv, q, vals, inds, None, 20., a), flush)
        dc = 1 if step == 0 else 3
        flops = 6561 * 98 * dc * 3
        tf = flops * q.shape[0] / ms / 1e9
        out["dist_chnls_%d" % dc] = dict(queries=int(q.shape[0]), lattice_pixels=int(nset), ms=ms,
                                         Mqueries_per_s=q.shape[0] / ms / 1e3, algorithmic_tflops=tf,
                                         frac_of_fp32_peak=tf / fp32_peak_tflops,
                                         no_reuse_window_GBps=(43.6e3 * dc + 1.2e3) * q.shape[0] / ms / 1e6)
    return out


def bayes_microbench(dev="cuda:0", ng=16384, fp32_peak_tflops=74.4, hbm_gbs=6547.2, flush=None):
    import vnlb_b200
    from vnlb_b200 import deno
    from vnlb_b200.utils import AttrDict
    flush = flush if flush is not None else torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}
    for step in (0, 1):
        n = 100 if step == 0 else 60
        a = vnlb_b200.get_args(vnlb_b200.get_params(20.), 3, step, dev)
        g = torch.Generator(device=dev)
        g.manual_seed(1)
        basis = torch.randn((ng, 3, 10, 98), device=dev, generator=g)
        coef = torch.randn((ng, 3, n, 10), device=dev, generator=g) * torch.linspace(40, 4, 10, device=dev)
        sig = torch.einsum("gcnr,gcrp->gcnp", coef, basis) / 3 + torch.rand((ng, 3, 1, 98), device=dev, generator=g) * 200
        noisy = sig + torch.randn(sig.shape, device=dev, generator=g) * 20
        basic = sig + torch.randn(sig.shape, device=dev, generator=g) * 3

        def unflat(x):
            return x.reshape(ng, 3, n, 2, 7, 7).permute(0, 2, 3, 1, 4, 5).contiguous()
        pn0, pb = unflat(noisy), unflat(basic if step == 1 else torch.zeros_like(noisy))
        del basis, coef, sig, noisy, basic
        patches = AttrDict(noisy=pn0.clone(), basic=pb, flat=torch.zeros(ng, dtype=torch.uint8, device=dev))
        ms = _timeit(lambda: deno.denoise(patches, a, "bayes"), flush)
        flops = 35.8e6 if step == 0 else 31.7e6
        bytes_ = (2 * n * 294 * 4) if step == 0 else (3 * n * 294 * 4)
        tf = flops * ng / ms / 1e9
        out["step%d" % (step + 1)] = dict(groups=ng, k=n, ms=ms, Mgroups_per_s=ng / ms / 1e3, nominal_tflops=tf,
                                          frac_of_fp32_peak=tf / fp32_peak_tflops, algorithmic_GBps=bytes_ * ng / ms / 1e6,
                                          frac_of_hbm_peak=bytes_ * ng / ms / 1e6 / hbm_gbs,
                                          one_million_groups_s=ms * 1e3 / ng)
        del patches, pn0, pb
    return out


if __name__ == "__main__":
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    ng = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    print(json.dumps(dict(search_config3=search_microbench(nq=nq), bayes_config4=bayes_microbench(ng=ng)), indent=1))

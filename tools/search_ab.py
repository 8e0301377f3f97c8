"""A/B of the search kernels (vnlb_set_search_path: 1 = 1-column tiled kernel, 2 = quad kernel + cp.async staging,
0 = quad kernel + TMA staging) on 960x540x16: BASELINE configs[2] (+-4 frames, no flow) and the production temporal
range (+-6 frames) with and without flows.  CUDA events, L2 flushed between iterations.
    python tools/search_ab.py [nq]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tools.microbench import _timeit


def main(nq=16384, dev="cuda:0"):
    import vnlb_b200
    from vnlb_b200 import _lib as L
    from vnlb_b200 import color, search, synth
    from vnlb_b200 import mask as gmask
    from vnlb_b200.utils import AttrDict
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    T, H, W = 16, 540, 960
    clean, flows = synth.synth_video(T, H, W, 123, return_flows=True)
    yuv = color.rgb2yuv(torch.from_numpy(synth.add_noise(clean, 20.)).to(dev))
    dflows = AttrDict(fflow=torch.from_numpy(flows["fflow"]).to(dev), bflow=torch.from_numpy(flows["bflow"]).to(dev))
    out = {}
    ref = {}
    for path in (1, 2, 0):
        L.lib.vnlb_set_search_path(path)
        for nwt, fl in ((4, None), (6, None), (6, dflows)):
            params = vnlb_b200.get_params(20.)
            params["sizeSearchTimeFwd"] = [nwt, nwt]
            params["sizeSearchTimeBwd"] = [nwt, nwt]
            params["nSimilarPatches"] = [100, 100]
            for step in (0, 1):
                a = vnlb_b200.get_args(params, 3, step, dev)
                m, nset = gmask.init_mask(yuv.shape, a, dev)
                q = torch.nonzero(m)
                q = q[:: max(1, q.shape[0] // nq)][:nq].contiguous()
                vals = torch.empty((q.shape[0], 100), device=dev)
                inds = torch.empty((q.shape[0], 100), dtype=torch.int64, device=dev)
                ms = _timeit(lambda: search.exec_sim_search_burst(yuv, q, vals, inds, fl, 20., a), flush)
                dc = 1 if step == 0 else 3
                ncand = (2 * nwt + 1) * 729
                tf = ncand * 98 * dc * 3 * q.shape[0] / ms / 1e9
                key = "nwt%d_%s_dc%d" % (nwt, "flow" if fl is not None else "noflow", dc)
                out.setdefault(key, {})["path%d" % path] = dict(ms=round(ms, 3), Mq_s=round(q.shape[0] / ms / 1e3, 3),
                                                                algorithmic_tflops=round(tf, 2))
                sig = (vals.double().sum().item(), int(inds.sum().item()))
                if key in ref:
                    out[key]["path%d" % path]["same_as_path1"] = sig == ref[key]
                else:
                    ref[key] = sig
    L.lib.vnlb_set_search_path(0)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 16384)

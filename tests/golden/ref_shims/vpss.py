"""Stand-in for the absent third-party `vpss` package, used ONLY by
tests/golden/make_golden.py: routes the reference's two call sites
(lib/vnlb/search/search.py:88-89,98) to the CPU oracle restatement."""
import numpy as np
import torch
from types import SimpleNamespace

from oracle import vnlb_oracle as orc


def _args(args):
    return SimpleNamespace(ps=args.ps, pt=args.pt, w_s=args.w_s, nWt_f=args.nWt_f, nWt_b=args.nWt_b,
                           npatches=args.npatches, step=args.step, c=args.c)


def exec_sim_search_burst(srch_img, srch_inds, vals, inds, flows, sigma, args):
    v = vals.numpy()
    i = inds.numpy()
    fl = {"fflow": flows.fflow.numpy(), "bflow": flows.bflow.numpy()}
    orc.exec_sim_search_burst(srch_img.numpy(), srch_inds.numpy(), v, i, fl, sigma, _args(args))


def fill_patches(patches, img, inds):
    orc.fill_patches(patches.numpy(), img.numpy(), inds.numpy())

"""Parameter table of the two VNLB steps and the per-step argument view.

Mirrors lib/vnlb/params.py of the reference: `default_params` :11-50 (classic
VNLB values: 7x7x2 patches, 27x27 window, +-6 frames, k=100/60, rank 39,
thresholds 2.7/0.7), `get_params` :52-100 (named overrides) and
`get_args`/`VnlbArgs` :102-214 (attribute view with the shortcuts ps, pt,
npatches, w_s, nWt_f, nWt_b, sigma2, sigmab2, thresh, patch_shape, bufs_shape).

Difference from the reference, on purpose: `get_params` there hard-codes
`version = "iphone"` (params.py:58), whose first step uses vpss's "needle"
search (not part of this hot path).  Here the version is an explicit argument
and defaults to the classic table the BASELINE configurations name."""
from ._lib import WINDOW_SHIFT, WINDOW_CLIP, EIG_TRIDIAG, EIG_JACOBI


def default_params(sigma, verbose=False):
    """lib/vnlb/params.py:11-50.  Every value is a [step1, step2] pair."""
    sigma = float(sigma)
    return dict(
        aggreBoost=[True, True], beta=[1.0, 1.0], bsize=[128, 128], c=[3, 3],
        coupleChannels=[False, False], device=["cpu", "cpu"], flatAreas=[False, True],
        gamma=[0.95, 0.2], isFirstStep=[True, False], mod_sel=["clipped", "clipped"],
        nParts=[-1, -1], nThreads=[-1, -1], nSimilarPatches=[100, 60], nkeep=[-1, -1],
        nstreams=[8, 18], offset=[2 * (sigma / 255.) ** 2, 0.], onlyFrame=[-1, -1],
        procStep=[3, 3], rank=[39, 39], sigma=[sigma, sigma], sigmaBasic=[sigma, 0.],
        sizePatch=[7, 7], sizePatchTime=[2, 2], sizeSearchTimeBwd=[6, 6],
        sizeSearchTimeFwd=[6, 6], sizeSearchWindow=[27, 27], step=[0, 1], tau=[0, 400.],
        testing=[False, False], use_imread=[False, False], stype=["l2", "l2"],
        srch_img=["noisy", "basic"], cpatches=["noisy", "basic"], deno=["bayes", "bayes"],
        var_mode=[0, 0], variThres=[2.7, 0.7], verbose=[verbose, verbose],
        # additions of this implementation (search semantics are runtime options)
        window_mode=["shift", "shift"], dist_chnls=[1, None], eig_method=["tridiag", "tridiag"],
    )


def get_params(sigma, verbose=False, version="default"):
    """lib/vnlb/params.py:52-100 with the version as an argument.  NOTE the default: the classic VNLB table of
    `default_params` (27x27 window, +-6 frames, 7x7x2 patches -- the values BASELINE.json names), whereas the
    reference's get_params is hard-coded to version = "iphone" (15x15 window, +-10 frames, pt = [1, 2], "needle" search
    in step 1).  Pass version="iphone" for that table (with the l2 search, and a warning)."""
    params = default_params(sigma, verbose)
    if version in ("default", "exp"):
        pass
    elif version in ("sss", "sss_v2", "iphone"):
        params["sizeSearchTimeBwd"] = [10, 10]
        params["sizeSearchTimeFwd"] = [10, 10]
        params["sizeSearchWindow"] = [15, 15]
        if version in ("sss_v2", "iphone"):
            params["sizePatchTime"] = [1, 2]
        if version == "iphone":
            # the reference pairs this with stype "needle" in step 1 (params.py:88), a
            # vpss search variant outside this path; the l2 search is used instead
            import warnings
            warnings.warn("vnlb_b200.get_params(version='iphone'): the reference's step-1 'needle' search (vpss) is not part of "
                          "this path; the l2 search is used for both steps", stacklevel=2)
            params["stype"] = ["l2", "l2"]
    else:
        raise ValueError("unknown params version [%s]" % version)
    return params


class VnlbArgs(dict):
    """Per-step view: attribute access + the reference's shortcuts (params.py:109-197)."""

    def __init__(self, params, step, c, device):
        super().__init__()
        for k, v in params.items():
            self[k] = v[step]
        self["step"] = step
        self["c"] = c
        self["device"] = device
        if self.get("dist_chnls") is None:
            self["dist_chnls"] = c

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    ps = property(lambda s: s["sizePatch"])
    ps_t = property(lambda s: s["sizePatchTime"])
    pt = property(lambda s: s["sizePatchTime"])
    npatches = property(lambda s: s["nSimilarPatches"])
    w_s = property(lambda s: s["sizeSearchWindow"])
    nWt_f = property(lambda s: s["sizeSearchTimeFwd"])
    nWt_b = property(lambda s: s["sizeSearchTimeBwd"])
    couple_ch = property(lambda s: s["coupleChannels"])
    sigma2 = property(lambda s: s["sigma"] ** 2)
    sigmaBasic2 = property(lambda s: s["sigmaBasic"] ** 2)
    sigmab2 = property(lambda s: s["sigmaBasic"] ** 2)
    thresh = property(lambda s: s["variThres"])
    flat_areas = property(lambda s: s["flatAreas"])
    step_s = property(lambda s: s["procStep"])
    tsize = property(lambda s: s["nstreams"] * s["bsize"])

    @property
    def patch_shape(self):                                   # params.py:222-227
        return (self.tsize, self.npatches, self.pt, self["c"], self.ps, self.ps)

    @property
    def bufs_shape(self):                                    # params.py:229-233
        return (self.tsize, self.npatches)

    @property
    def window_mode_id(self):
        return {"shift": WINDOW_SHIFT, "clip": WINDOW_CLIP}[self["window_mode"]]

    @property
    def eig_method_id(self):
        return {"tridiag": EIG_TRIDIAG, "jacobi": EIG_JACOBI}[self["eig_method"]]


def get_args(params, c, step, device):
    """lib/vnlb/params.py:102-214."""
    if step not in (0, 1):
        raise ValueError("step must be 0 or 1")
    for key in ("sizeSearchWindow",):
        if params[key][step] % 2 == 0:
            raise ValueError("sizeSearchWindow must be odd")
    return VnlbArgs(params, step, c, device)

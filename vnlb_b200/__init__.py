"""vnlb_b200 -- Blackwell-native Video Non-Local Bayes.

Drop-in for the `vnlb` package's denoise path: `vnlb_b200.denoise(noisy, sigma,
flows=...) -> (deno, basic, dtime)`.  PyTorch is plumbing (device memory,
streams, torch.distributed); every stage runs in hand-written sm_100a CUDA
kernels behind the C ABI of include/vnlb_b200.h.  No CPU fallback."""
from . import _lib  # noqa: F401  (fails loudly if libvnlb_b200.so is missing)
from .impl import denoise
from .params import default_params, get_args, get_params
from .proc_nl import proc_nl
from .utils import compute_psnrs

__all__ = ["denoise", "default_params", "get_params", "get_args", "proc_nl", "compute_psnrs"]

"""Smallest reproduction of a search call (debugging aid): python tools/dbg_search.py [path] [W]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import vnlb_b200
from vnlb_b200 import _lib as L
from vnlb_b200 import search

path = int(sys.argv[1]) if len(sys.argv) > 1 else 0
W = int(sys.argv[2]) if len(sys.argv) > 2 else 64
L.lib.vnlb_set_search_path(path)
T, H = 3, 64
img = torch.from_numpy((np.random.RandomState(0).rand(T, 3, H, W) * 255).astype(np.float32)).cuda()
a = vnlb_b200.get_args(vnlb_b200.get_params(20.), 3, 0, "cuda:0")
q = torch.tensor([[0, 0, 0], [1, 20, 30], [1, H - 7, W - 7]], dtype=torch.int64, device="cuda:0")
vals = torch.empty((3, a.npatches), device="cuda:0")
inds = torch.empty((3, a.npatches), dtype=torch.int64, device="cuda:0")
search.exec_sim_search_burst(img, q, vals, inds, None, 20., a)
torch.cuda.synchronize()
print("path", path, "ok", vals[:, :3].tolist(), inds[:, :3].tolist())

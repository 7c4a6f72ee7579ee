import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from photogrammetry_b200.keypoint_matching import Matcher
m = Matcher(0); n, dim = int(os.environ.get("L2_N", "16384")), 128
q = torch.rand((n, dim), device="cuda"); t = torch.rand((n, dim), device="cuda")
oj = torch.empty((2, n), dtype=torch.int32, device="cuda"); od = torch.empty((2, n), device="cuda")
for _ in range(3):
    m._check(m._lib.pgm_knn2_l2_dev(m._h, q.data_ptr(), n, t.data_ptr(), n, dim, oj[0].data_ptr(), od[0].data_ptr(), oj[1].data_ptr(), od[1].data_ptr(), None))
m.synchronize(); print("ok")

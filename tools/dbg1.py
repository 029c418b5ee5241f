import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/2d-lbm-dem_b200"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import lbmdem_gpu as G
from util import small_packing
k = int(sys.argv[1]); prec = sys.argv[2]
s = G.Solver(70, 131, 1.0, prec, kernel=k, strict_fp=int(sys.argv[3]))
r, x, y = small_packing(70, 131, 1.0, 11)
s.init_arrays(r, x, y)
s.lbm_step()
print("ok", k, prec, s.total_density())

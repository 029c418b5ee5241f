"""tools/default_build_library.py -- the reference's DEFAULT build (7826 x 2325 fp64) on bin/50000-test.data through the
library alone (lbmdem_step, state resident): ms per coupled step, batched and one renderScene() call at a time
(profiles/r02_default_build_application.txt)."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT+"/2d-lbm-dem_b200"): sys.path.insert(0,p)
import lbmdem_gpu as G, torch
s = G.Solver(7826, 2325, 1.0, "f64")
n = s.init(ROOT+"/tests/golden/50000-test.data")
npd = s.scalars()["npDEM"]; print("n", n, "npDEM", npd)
s.step(npd*10); torch.cuda.synchronize()
t0=time.time(); s.step(npd*50); torch.cuda.synchronize(); t1=time.time()
print("batched: %.3f ms per coupled step" % ((t1-t0)/50*1e3), s.list_counts())
t0=time.time()
for k in range(npd*50): s.step(1)
torch.cuda.synchronize(); t1=time.time()
print("one call at a time: %.3f ms per coupled step" % ((t1-t0)/50*1e3))

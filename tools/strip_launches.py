"""tools/strip_launches.py -- the kernel sequence of the STRIP path in one process, for a launch list under ncu.

ncu must not wrap a multi-rank job; the in-process strip group (lbmdem_local_group_*) runs the same per-rank kernel
sequence as the NCCL path -- fused kernel on the owned rows, interior sweeps, ghost rows, edge sweeps, force links, sum
over the ranks, DEM -- with peer copies and one summing kernel in place of ncclSend/Recv and ncclAllReduce.  Two strips
of 4096 rows of the benchmarked sample (BASELINE configs[3] repeated twice along x) on device 0.

    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file out.csv python tools/strip_launches.py
"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "2d-lbm-dem_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)

import bench  # noqa: E402
import lbmdem_dist as D  # noqa: E402


def main():
    strips = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    path = os.path.join(tempfile.mkdtemp(prefix="strip_launches_"), "s.data")
    n, what = bench.make_sample_file("cfg4", strips, path)
    grp = D.LocalStrips(4096 * strips, 4096, 2.7, "f32", strips)
    assert grp.init(path) == n
    npd = grp.ranks[0].scalars()["npDEM"]
    grp.step(npd * steps)
    print(f"{strips} strips of 4096 x 4096 fp32 on one device, {n} grains ({what}), {steps} coupled steps")
    grp.close()


if __name__ == "__main__":
    main()

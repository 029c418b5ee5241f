"""world_size-2 and -3 CPU runs (gloo) of the strip decomposition: ghost width, sweep ranges,
row exchange and the integer all-reduce of the force sums, with the product's node headers
(tests/hostcheck) standing in for the kernels.  See tests/strip_worker.py."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_strip_decomposition_matches_the_oracle_and_is_decomposition_independent(world):
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "strip_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {rank} failed:\n{out[-3000:]}"
        assert f"rank {rank}: ok" in out


def test_strip_bounds_cover_the_lattice():
    import lbmdem_dist as D
    for lx in (8, 61, 4096, 8190):
        for n in (1, 2, 3, 8):
            b = [D.strip_bounds(lx, r, n) for r in range(n)]
            assert b[0][0] == 0 and b[-1][1] == lx
            assert all(b[k][1] == b[k + 1][0] for k in range(n - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1

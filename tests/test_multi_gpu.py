"""Strip decomposition on real GPUs (needs >= 2 visible devices): one process per GPU over NCCL."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("peer_sums", ["1", "0"])
def test_strips_are_bit_identical_to_one_gpu(peer_sums):
    """peer_sums: the force sums of the ranks added through CUDA IPC mappings of the peers' buffers (the default) or
    with ncclAllReduce (LBMDEM_PEER_SUMS=0); both must reproduce the one-GPU run bit for bit."""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least two GPUs")
    world = min(ngpu, 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(HERE, "multigpu_worker.py")]
    env = dict(os.environ, LBMDEM_PEER_SUMS=peer_sums, LBMDEM_VERBOSE="1")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, (p.stdout + p.stderr)[-4000:]
    for rank in range(world):
        assert f"rank {rank}: ok" in p.stdout
    assert ("peer-memory force sums" in p.stderr) == (peer_sums == "1"), p.stderr[-2000:]
    if peer_sums == "1":
        assert "peer-memory force sums on" in p.stderr, "CUDA IPC between the ranks' GPUs is expected to work on one box"


def test_executable_on_two_gpus_writes_the_same_files_as_on_one(tmp_path):
    """`lbmdem <file> --gpus 2` (the executable forks one rank per GPU): the default build is bit-identical for any
    number of strips, so every output file must equal the one-GPU run's byte for byte."""
    import importlib.util
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    root = os.path.dirname(HERE)
    spec = importlib.util.spec_from_file_location("host_build", os.path.join(root, "2d-lbm-dem_b200", "host", "build.py"))
    hb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(hb)
    exe = hb.build_exe()
    sample = os.path.join(HERE, "golden", "pack_64x48_f64.data")
    outs = []
    for gpus in (1, 2):
        out = tmp_path / f"g{gpus}"
        out.mkdir()
        p = subprocess.run([exe, sample, "--lx", "64", "--ly", "48", "--steps", "8000", "--gpus", str(gpus), "--outdir", str(out)],
                           capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
        assert "final_density:" in p.stderr
        outs.append((out, p))
    names = sorted(os.listdir(outs[0][0]))
    assert len(names) == 8 and names == sorted(os.listdir(outs[1][0]))
    for name in names:
        assert open(outs[0][0] / name, "rb").read() == open(outs[1][0] / name, "rb").read(), name
    assert outs[0][1].stderr.strip().splitlines()[-1] == outs[1][1].stderr.strip().splitlines()[-1]   # final_density

"""The NS exchange over NVLink peer memory (HANS_NS_PEER=1, hans_ns_exchange_peer) against the NCCL all_gather
path: same volumes.txt, line for line, on two GPUs.  Needs two devices; skipped on a one-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

INPUT = """move_ratio = 3,48,32,16,3,3
nbeads = 6
nchains = 16
nwalkers = 8
ranks_per_gpu = 16
iterations = 300
walklength = 5
initial_walk = 10
seed = 5
directory = {d}
"""


def _run(tmp_path, name, peer, port):
    inp = tmp_path / f"{name}.txt"
    inp.write_text(INPUT.format(d=name))
    env = dict(os.environ, HANS_NS_PEER="1" if peer else "0", PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), "-m", "hans_b200.mpihans", "-f", str(inp)]
    r = subprocess.run(cmd, cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return (tmp_path / name / "volumes.txt").read_text().splitlines()


def test_peer_exchange_equals_nccl(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    a = _run(tmp_path, "nccl", False, 29641)
    b = _run(tmp_path, "peer", True, 29642)
    assert len(a) == 301 and a == b

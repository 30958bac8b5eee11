"""The NS exchange over NVLink peer memory (HANS_NS_PEER=1, hans_ns_exchange_peer): against the NCCL all_gather
path on two GPUs (same volumes.txt, line for line; skipped on a one-GPU box), and its protocol with two ranks
emulated on one GPU (passes on B200)."""
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


def _run(tmp_path, name, peer, port, nproc=2, rpg=16):
    inp = tmp_path / f"{name}.txt"
    inp.write_text(INPUT.format(d=name).replace("ranks_per_gpu = 16", f"ranks_per_gpu = {rpg}"))
    env = dict(os.environ, HANS_NS_PEER="1" if peer else "0", PYTHONPATH=ROOT)
    if nproc > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
               "127.0.0.1", "--master-port", str(port), "-m", "hans_b200.mpihans", "-f", str(inp)]
    else:
        cmd = [sys.executable, "-m", "hans_b200.mpihans", "-f", str(inp)]
    r = subprocess.run(cmd, cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return (tmp_path / name / "volumes.txt").read_text().splitlines()


def test_two_gpu_run_equals_one_gpu_run(tmp_path):
    """The same 32 virtual ranks x 8 walkers laid out as 2 GPUs x 16 (NCCL all_gather exchange) and as 1 GPU x 32 write the
    same volumes.txt, line for line, and the same restart file (groups are named per virtual rank): walker ids, clone
    source, MAXLOC tie rule, active walkers and the step adaptation are all functions of GLOBAL indices
    (/root/reference/mpihans.py:215-245: the result of the reference does not depend on where a rank runs either)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import numpy as np
    from hans_b200 import NSio
    two = _run(tmp_path, "two", False, 29651, nproc=2, rpg=16)
    one = _run(tmp_path, "one", False, 29652, nproc=1, rpg=32)
    assert len(two) == 301 and two == one
    a1, w1 = NSio.open_restart(str(tmp_path / "one" / "restart.hdf5"))
    a2, w2 = NSio.open_restart(str(tmp_path / "two" / "restart.hdf5"))
    assert sorted(w1) == sorted(w2) and len(w1) == 256
    for g in w1:
        assert np.array_equal(w1[g][0], w2[g][0]) and np.array_equal(w1[g][1], w2[g][1]) and w1[g][2] == w2[g][2], g
    for k in ("dv_max", "dr_max", "dt_max", "dh_max", "dshear", "dstretch", "prev_iters"):
        assert a1[k] == a2[k], k


def test_peer_exchange_equals_nccl(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    a = _run(tmp_path, "nccl", False, 29641)
    b = _run(tmp_path, "peer", True, 29642)
    assert len(a) == 301 and a == b


@pytest.mark.skipif(os.environ.get("HANS_TEST_PEER", "0") != "1",
                    reason="opt-in (HANS_TEST_PEER=1): two kernels on two streams of one GPU must be co-resident, which CUDA does not "
                           "promise; the real test is test_peer_exchange_equals_nccl on two GPUs")
def test_peer_protocol_two_ranks_emulated_on_one_gpu():
    """The protocol of hans_ns_exchange_peer without IPC: two 'ranks' on ONE device, their buffers plain
    allocations of that device, their kernels on two streams (they must be co-resident: one CTA each).  Decisions
    must equal hans_ns_pack + hand-made gather + hans_ns_resolve on copies of the same pools.  Waits are bounded
    (hans_ns_peer_timeout) so that a protocol error fails the test instead of hanging the GPU."""
    import ctypes as C
    import numpy as np
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from hans_b200 import _lib, engine
    from hans_b200.ns import NestedSampler
    L = _lib.lib
    nl, mr = 6, [1, 18, 12, 0, 3, 3]

    def pools():
        ps = []
        for r in range(2):
            p = engine.WalkerPool(nl, 6, 2, mr, seed=3, gid_base=r * nl)
            p.set_initial_cells(); p.grow(); p.mc_run(2 + 3 * r, counters=False)
            ps.append(p)
        return ps

    ref, new = pools(), pools()
    s_ref = [NestedSampler(ref[r], nwalkers=nl, walklength=1, rank=r, world=1, seed=5, traj_interval=0) for r in range(2)]
    s_new = [NestedSampler(new[r], nwalkers=nl, walklength=1, rank=r, world=1, seed=5, traj_interval=0) for r in range(2)]
    rec_len = s_ref[0].rec_len
    gathered = torch.zeros(2 * rec_len, dtype=torch.float64, device=ref[0].device)
    for r in range(2):
        for s in (s_ref[r], s_new[r]):
            s.args.world, s.args.rank, s.args.d_gathered = 2, r, gathered.data_ptr()
    nbytes = int(L.hans_ns_peer_bytes(ref[0].N, 2))
    bufs = [torch.zeros(nbytes // 8, dtype=torch.int64, device=ref[0].device) for _ in range(2)]
    table = (C.c_void_p * 2)(bufs[0].data_ptr(), bufs[1].data_ptr())
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    _lib.check(L.hans_ns_peer_timeout(5.0))
    _lib.diag_counters(reset=True)
    st = torch.cuda.current_stream().cuda_stream
    try:
        for it in range(10):
            for s in s_ref:
                _lib.check(L.hans_ns_pack(C.byref(s.args), st))
            gathered[:rec_len].copy_(s_ref[0].rec)
            gathered[rec_len:].copy_(s_ref[1].rec)
            for s in s_ref:
                _lib.check(L.hans_ns_resolve(C.byref(s.args), st))
            torch.cuda.synchronize()
            for r in (1, 0):  # rank 1 first: it has to wait for rank 0's kernel on the other stream
                _lib.check(L.hans_ns_exchange_peer(C.byref(s_new[r].args), table, streams[r].cuda_stream))
            torch.cuda.synchronize()
            assert _lib.diag_counters()[2] == 0, "a wait for the other rank's flag expired"
            for r in range(2):
                assert s_new[r].volume_limit() == s_ref[r].volume_limit()
                for a, b in zip(new[r].get_state(), ref[r].get_state()):
                    assert np.array_equal(a, b)
                assert np.array_equal(new[r].vol.cpu().numpy(), ref[r].vol.cpu().numpy())
                assert np.array_equal(s_new[r].d_ids.cpu().numpy(), s_ref[r].d_ids.cpu().numpy())
                assert int(s_new[r].d_iter.item()) == it + 1
            for r in range(2):  # the same walk on both copies so that volumes change between iterations
                for p, s in ((ref[r], s_ref[r]), (new[r], s_new[r])):
                    p.mc_run(1, ids=s.d_ids, d_volume_limit=s.d_limit, counters=False)
            torch.cuda.synchronize()
    finally:
        L.hans_ns_peer_timeout(0.0)

"""world_size = 2 on CPU (gloo): the multi-rank host logic of the NS loop -- record exchange,
identical MAXLOC / clone decisions on every rank, rank-0-only logging, and the cross-rank
step-size averaging of adjust_mc_steps (MCNS.py:685)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from hans_b200 import MCNS as NS
        from hans_b200 import mpihans, philox
        from hans_b200.ns import resolve_host, update_steps

        # --- adjust_mc_steps averaging: Allreduce(SUM)/size, same update on every rank ---------------
        comm = NS.TorchComm()
        assert comm.Get_size() == world and comm.Get_rank() == rank
        rate = np.array([0.1, 0.9, 0.3, 0.0, 0.6, 0.1]) if rank == 0 else np.array([0.2, 0.5, 0.3, 0.0, 0.8, 0.2])
        avg = np.zeros(6)
        comm.Allreduce(rate, avg)
        avg /= world
        assert np.allclose(avg, [0.15, 0.7, 0.3, 0.0, 0.7, 0.15])
        steps = dict(dv_max=2.0, dr_max=2.0, dt_max=0.43, dh_max=0.4, dshear=1.0, dstretch=1.0)
        new = update_steps(steps, avg, [3, 48, 32, 0, 3, 3])
        assert new == dict(dv_max=1.0, dr_max=4.0, dt_max=0.43, dh_max=0.4, dshear=2.0, dstretch=0.5)

        # --- the sharded NS bookkeeping (host model of hans_ns_pack / all-gather / hans_ns_resolve) ---
        nlocal, N, seed = 6, 4, 99
        rng = np.random.default_rng(1000 + rank)
        vol = rng.uniform(50, 100, nlocal)
        R = rng.normal(size=(nlocal, N, 3))
        H = np.tile(np.eye(3), (nlocal, 1, 1)) * np.cbrt(vol)[:, None, None]
        log = []
        fvol = open(os.path.join(tmp, "volumes.txt"), "a+") if rank == 0 else None
        if rank == 0:
            fvol.write(mpihans.volumes_header(nlocal * world, 0, N))
        for it in range(60):
            sr, si = philox.clone_source(seed, it, nlocal, world)
            rec = np.zeros(12 + 3 * N)
            rec[0], rec[1] = vol.max(), int(np.argmax(vol))                     # mpihans.py:211-212
            if sr == rank:                                                      # mpihans.py:226-229
                rec[2], rec[3:12], rec[12:] = vol[si], H[si].ravel(), R[si].ravel()
            gathered = [torch.zeros(12 + 3 * N, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(gathered, torch.from_numpy(rec))                    # one collective per iteration
            G = np.stack([g.numpy() for g in gathered])
            vmax, mr, mi, sr2, si2 = resolve_host(G, seed, it, nlocal)
            assert (sr2, si2) == (sr, si)
            everyone = [None] * world
            dist.all_gather_object(everyone, vol.copy())
            allv = np.concatenate(everyone)
            assert vmax == allv.max() and mr * nlocal + mi == int(np.argmax(allv))   # global MAXLOC, lowest index on ties
            if rank == mr:                                                      # mpihans.py:232-237
                vol[mi], H[mi], R[mi] = G[sr, 2], G[sr, 3:12].reshape(3, 3), G[sr, 12:].reshape(N, 3)
                assert vol[mi] == everyone[sr][si]
            active = mi if rank == mr else int(philox.ns_draw(seed, it, rank, 1) % nlocal)   # mpihans.py:236-239
            vol[active] *= rng.uniform(0.9, 1.0)                                 # stands in for the walk under vmax
            assert vol[active] <= vmax
            log.append(vmax)
            if rank == 0:
                fvol.write(mpihans.volumes_line(it, vmax))                       # rank 0 only, mpihans.py:219-221
        # every rank saw the same monotone ceiling sequence
        logs = [None] * world
        dist.all_gather_object(logs, log)
        assert logs[0] == logs[1] and (np.diff(log) <= 0).all()
        if rank == 0:
            fvol.close()
        dist.barrier()
        lines = open(os.path.join(tmp, "volumes.txt")).read().splitlines()
        assert len(lines) == 61 and lines[0] == "12 1 0 False 4 "
    finally:
        dist.destroy_process_group()


def test_two_rank_ns_protocol_on_gloo(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)

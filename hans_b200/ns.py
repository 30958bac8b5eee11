"""Nested-sampling iteration engine on top of WalkerPool (host side of mpihans.py:210-281).

One process per GPU (rank == GPU); walkers are sharded `nlocal` per rank exactly as the reference
shards `nwalkers` per MPI rank (mpihans.py:54,185).  A GPU may host several *virtual ranks* of
`nwalkers` walkers each; every virtual rank behaves like one MPI rank of the reference: per
iteration it walks one walker -- the freshly cloned max-volume slot if it owns it, a uniformly
random one of its walkers otherwise (mpihans.py:232-241, pymatnest-style parallel decorrelation).

Per iteration the host enqueues, with no synchronisation:
    hans_ns_pack     local MAXLOC + staging of the clone source            (mpihans.py:211-212,226-229)
    all_gather       one NCCL collective of world x (12+3N) doubles        (mpihans.py:215,223,229)
    hans_ns_resolve  global MAXLOC, ceiling, log, clone, active walkers    (mpihans.py:215-239)
    hans_mc_run_batch  walklength sweeps for the active walkers            (mpihans.py:241-245)
The clone source is a shared Philox draw keyed by the iteration number, so no broadcast of the
choice is needed; the ceiling, iteration counter and active list stay in device memory.

With HANS_NS_PEER=1 (one node, world > 1) the first three become ONE kernel, hans_ns_exchange_peer:
records are stored straight into CUDA-IPC-mapped buffers of the other ranks over NVLink and a flag
per rank and iteration replaces the collective (include/hans_b200.h).  Verified against the NCCL path on two B200s and
timed at eight (equal: the exchange is ~1 % of a step); NCCL stays the default.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

from . import _lib, philox
from ._lib import NsArgs, check, lib
from .engine import WalkerPool, _stream

# mpihans.py:24-33
STEP_MAX = dict(dv_max=50.0, dr_max=50.0)
STEP_MIN = np.array([1e-10, 1e-10, 1e-10, 1e-10, 1e-5, 1e-5])
STEP_NAMES = ["dv_max", "dr_max", "dt_max", "dh_max", "dshear", "dstretch"]


def update_steps(steps: dict, avg_rate, move_ratio, lower_bound=0.2, upper_bound=0.5, min_dstep=STEP_MIN,
                 dv_max=STEP_MAX["dv_max"], dr_max=STEP_MAX["dr_max"]) -> dict:
    """The step-size rule of adjust_mc_steps (MCNS.py:687-716) as a pure function.

    < lower_bound -> halve (floored at min_dstep); > upper_bound -> double (dv, dr capped at
    dv_max / dr_max; the others uncapped, as in the reference)."""
    out = dict(steps)
    caps = [dv_max, dr_max, None, None, None, None]
    for i, name in enumerate(STEP_NAMES):
        if move_ratio[i] == 0:
            continue
        if avg_rate[i] < lower_bound:
            out[name] = max(0.5 * steps[name], float(min_dstep[i]))
        elif avg_rate[i] > upper_bound:
            out[name] = 2.0 * steps[name] if caps[i] is None else min(2.0 * steps[name], caps[i])
    return out


def resolve_host(records: np.ndarray, seed: int, iteration: int, nlocal: int):
    """Numpy statement of hans_ns_resolve's decisions for records [world, 12+3N] (used by tests and
    the gloo multi-rank tests): returns (vmax, max_rank, max_idx, src_rank, src_idx)."""
    world = records.shape[0]
    best, br, bi = -np.inf, 0, 0
    for r in range(world):
        if records[r, 0] > best:
            best, br, bi = records[r, 0], r, int(records[r, 1])
    sr, si = philox.clone_source(seed, iteration, nlocal, world)
    return float(best), br, bi, int(sr), int(si)


class NestedSampler:
    """Device-resident NS loop for one rank.

    pool          WalkerPool with nlocal = nwalkers * vranks walkers
    nwalkers      walkers per virtual rank (the reference's `nwalkers`, per MPI rank)
    walklength    sweeps per walk (mpihans.py:241)
    group         torch.distributed process group (None when world == 1)
    """

    def __init__(self, pool: WalkerPool, nwalkers: int, walklength: int, rank: int = 0, world: int = 1,
                 group=None, seed: int = 0, start_iter: int = 0, traj_interval: int = 100, traj_slots: int = 8,
                 log_cap: int = 8192, mc_adjust_wl: int = 10, adjust_samples: int = 1024):
        if pool.nwalkers % nwalkers:
            raise ValueError("pool size must be a multiple of nwalkers")
        self.pool, self.nw, self.walklength = pool, int(nwalkers), int(walklength)
        self.rank, self.world, self.group = int(rank), int(world), group
        self.seed = int(seed)
        self.nvranks = pool.nwalkers // self.nw
        self.K = pool.nwalkers * self.world  # total live walkers (mpihans.py:202)
        self.mc_adjust_wl = int(mc_adjust_wl)
        self.adjust_samples = int(adjust_samples)  # virtual ranks sampled per adaptation, over ALL GPUs
        dev = pool.device
        self.rec_len = 12 + 3 * pool.N
        self.rec = torch.zeros(self.rec_len, dtype=torch.float64, device=dev)
        self.gathered = self.rec if self.world == 1 else torch.zeros(self.world * self.rec_len, dtype=torch.float64, device=dev)
        self.d_iter = torch.full((1,), int(start_iter), dtype=torch.int64, device=dev)
        self.d_limit = torch.zeros(1, dtype=torch.float64, device=dev)
        self.log_cap = int(log_cap)
        self.d_log = torch.zeros((self.log_cap, 3), dtype=torch.float64, device=dev)
        self.d_ids = torch.zeros(self.nvranks, dtype=torch.int32, device=dev)
        self.traj_interval, self.traj_slots = int(traj_interval), int(traj_slots)
        self.d_traj = torch.zeros((self.traj_slots, self.rec_len), dtype=torch.float64, device=dev)
        self.d_traj_iter = torch.full((self.traj_slots,), -1, dtype=torch.int64, device=dev)
        self.iteration = int(start_iter)  # host mirror of d_iter
        self.flushed = int(start_iter)
        self.walk_counters = False
        self.att = self.acc = None
        self._scratch: Optional[WalkerPool] = None
        a = NsArgs()
        a.N, a.nlocal, a.rank, a.world = pool.N, pool.nwalkers, self.rank, self.world
        a.nw_per_vrank, a.traj_interval, a.traj_slots, a.log_cap = self.nw, self.traj_interval, self.traj_slots, self.log_cap
        a.seed = self.seed
        a.d_H, a.d_R, a.d_vol = pool.H.data_ptr(), pool.R.data_ptr(), pool.vol.data_ptr()
        a.d_iter, a.d_rec, a.d_gathered = self.d_iter.data_ptr(), self.rec.data_ptr(), self.gathered.data_ptr()
        a.d_limit, a.d_log, a.d_ids = self.d_limit.data_ptr(), self.d_log.data_ptr(), self.d_ids.data_ptr()
        a.d_traj, a.d_traj_iter = self.d_traj.data_ptr(), self.d_traj_iter.data_ptr()
        self.args = a
        self._peer = None
        if self.world > 1 and os.environ.get("HANS_NS_PEER", "0") == "1":
            self._setup_peer()

    def _setup_peer(self) -> None:
        """Map one exchange buffer per rank into every rank of the node (CUDA IPC); include/hans_b200.h."""
        import torch.distributed as dist
        if self.world > 16:
            raise ValueError("HANS_NS_PEER supports up to 16 ranks on one node")
        nbytes = int(lib.hans_ns_peer_bytes(self.pool.N, self.world))
        mine = C.c_void_p()
        handle = C.create_string_buffer(64)
        check(lib.hans_peer_alloc(nbytes, C.byref(mine), handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=self.group)
        table = (C.c_void_p * self.world)()
        opened = []
        for r in range(self.world):
            if r == self.rank:
                table[r] = mine.value
            else:
                p = C.c_void_p()
                check(lib.hans_peer_open(handles[r], C.byref(p)))
                table[r] = p.value
                opened.append(p.value)
        dist.barrier(group=self.group)
        self._peer = dict(mine=mine.value, opened=opened, table=table)

    def close(self) -> None:
        """Release the peer-memory exchange buffers (no-op on the NCCL path)."""
        if self._peer is not None:
            import torch.distributed as dist
            torch.cuda.synchronize(self.pool.device)
            dist.barrier(group=self.group)
            for p in self._peer["opened"]:
                check(lib.hans_peer_close(p))
            dist.barrier(group=self.group)
            check(lib.hans_peer_free(self._peer["mine"]))
            self._peer = None

    # ---- the iteration --------------------------------------------------------------------------
    def step_with_events(self, ev0, ev1) -> None:
        """step() with CUDA events recorded around the walk kernel (bench.py roofline timing)."""
        self.step(ev0, ev1)

    def step(self, ev0=None, ev1=None) -> None:
        """Enqueue one NS iteration (mpihans.py:210-245).  No host synchronisation."""
        st = _stream()
        if self._peer is not None:
            check(lib.hans_ns_exchange_peer(C.byref(self.args), self._peer["table"], st))
            self.pool.launches += 1
        else:
            check(lib.hans_ns_pack(C.byref(self.args), st))
            if self.world > 1:
                import torch.distributed as dist
                dist.all_gather_into_tensor(self.gathered, self.rec, group=self.group)
            check(lib.hans_ns_resolve(C.byref(self.args), st))
            self.pool.launches += 2
        ids = None if self.nw == 1 else self.d_ids  # nw == 1: every resident walker walks
        if ev0 is not None:
            ev0.record()
        self.att, self.acc = self.pool.mc_run(self.walklength, ids=ids, d_volume_limit=self.d_limit,
                                              counters=self.walk_counters)
        if ev1 is not None:
            ev1.record()
        self.iteration += 1

    def run(self, n: int) -> None:
        for _ in range(int(n)):
            self.step()

    @property
    def moves_per_iteration(self) -> int:
        """MC move attempts one iteration performs on this rank."""
        return self.nvranks * self.walklength * self.pool.moves_per_sweep

    # ---- results ---------------------------------------------------------------------------------
    def flush(self):
        """Synchronise and return what the iterations since the last flush produced:
        (log, frames): log = array [n,4] of (iteration, max volume, owner rank, owner index);
        frames = list of (iteration, H(3,3), R(N,3)) staged on THIS rank (old max walkers)."""
        n = self.iteration - self.flushed
        if n > self.log_cap:
            raise RuntimeError("volume log ring overflowed: flush more often than log_cap iterations")
        torch.cuda.synchronize(self.pool.device)
        if self._peer is not None and _lib.diag_counters()[2]:
            raise _lib.HansError("peer-memory NS exchange: a wait for another rank's flag expired (dead or stalled rank); "
                                 "the iterations since the last flush are void")
        its = np.arange(self.flushed, self.iteration, dtype=np.int64)
        log = np.zeros((n, 4))
        if n:
            dl = self.d_log.cpu().numpy()
            log[:, 0] = its
            log[:, 1:] = dl[its % self.log_cap]
        frames = []
        if self.traj_interval > 0:
            ti = self.d_traj_iter.cpu().numpy()
            hit = np.nonzero(ti >= self.flushed)[0]
            if hit.size:
                tr = self.d_traj.cpu().numpy()
                for s in hit[np.argsort(ti[hit])]:
                    frames.append((int(ti[s]), tr[s, 3:12].reshape(3, 3).copy(), tr[s, 12:].reshape(-1, 3).copy()))
                self.d_traj_iter.fill_(-1)
        self.flushed = self.iteration
        return log, frames

    def volume_limit(self) -> float:
        return float(self.d_limit.item())

    # ---- step-size adaptation (MCNS.py:657-717) --------------------------------------------------
    def steps(self) -> dict:
        c = self.pool.cfg
        return {k: float(getattr(c, k)) for k in STEP_NAMES}

    def set_steps(self, steps: dict) -> None:
        for k in STEP_NAMES:
            setattr(self.pool.cfg, k, float(steps[k]))
            if self._scratch is not None:
                setattr(self._scratch.cfg, k, float(steps[k]))

    def measure_rates(self, vol_max: Optional[float] = None) -> np.ndarray:
        """Acceptance rate of every enabled move type from `mc_adjust_wl` single-type sweeps on copies of one random
        walker of each sampled virtual rank under the current ceiling (MCNS.py:676-684; the copies replace the backup /
        import_ase_to_ibox restore).  Returns [sum of rates (6), number of sampled virtual ranks] for the cross-rank
        average.

        Everything random here is a pure function of (seed, iteration, GLOBAL virtual rank): which virtual ranks are
        sampled (at most `adjust_samples` of all world x nvranks), which of its walkers each one tries, and the Philox
        stream of every trial walk (walker id 0x40000000 + 6 x global virtual rank + move type, counter positioned by
        the iteration).  So a run resumed from a restart file, and a run laid out over another number of GPUs with
        the same total of virtual ranks, adapt exactly like the original (the reference saves no RNG state at all)."""
        pool, nv, nw = self.pool, self.nvranks, self.nw
        total = nv * self.world
        if self._scratch is None:
            self._scratch = WalkerPool(nv * 6, pool.nchains, pool.nbeads, pool.move_ratio, device=pool.device,
                                       seed=pool.seed, gid_base=0x40000000 + self.rank * nv * 6,
                                       threads_per_walker=pool.tpw, **pool._cfg_kw)
        sc = self._scratch
        C.memmove(C.byref(sc.cfg), C.byref(pool.cfg), C.sizeof(_lib.Cfg))
        rng = np.random.default_rng([self.seed, self.iteration, 0x5ca1ab1e])
        S = min(total, self.adjust_samples)
        chosen = rng.choice(total, size=S, replace=False) if S < total else np.arange(total)
        walker_of = rng.integers(0, nw, size=total)            # the walker each virtual rank would try (MCNS.py:676)
        lo = self.rank * nv
        mine = np.sort(chosen[(chosen >= lo) & (chosen < lo + nv)]) - lo
        sums = np.zeros(7)
        sums[6] = len(mine)
        if len(mine) == 0:
            return sums
        src = torch.as_tensor(np.repeat(mine * nw + walker_of[lo + mine], 6), device=pool.device)
        dst = torch.as_tensor((mine[:, None] * 6 + np.arange(6)[None, :]).reshape(-1), device=pool.device)
        sc.H[dst] = pool.H[src]
        sc.R[dst] = pool.R[src]
        sc.ctr.fill_(self.iteration << 24)
        limit = self.d_limit if vol_max is None else None
        outs = []
        for i in range(6):
            if pool.move_ratio[i] == 0:
                continue
            ratio = np.zeros(6)
            ratio[i] = 1.0  # np.eye(6)[i], MCNS.py:679-682
            ids = torch.as_tensor((mine * 6 + i).astype(np.int32), device=pool.device)
            att, acc = sc.mc_run(self.mc_adjust_wl, ids=ids, d_volume_limit=limit,
                                 volume_limit=_lib.DBL_MAX if vol_max is None else float(vol_max),
                                 cfg=sc.cfg_with(ratio), count_pairs=False)
            outs.append((i, att, acc))
        pool.launches += sc.launches
        sc.launches = 0
        for i, att, acc in outs:
            a = att[:, i].double().clamp(min=1.0)  # MCNS.py:384
            sums[i] = float((acc[:, i].double() / a).sum().item())
        return sums

    def adjust_mc_steps(self, vol_max: Optional[float] = None, lower_bound=0.2, upper_bound=0.5):
        """adjust_mc_steps (MCNS.py:657-717): returns the cross-rank average rates."""
        sums = self.measure_rates(vol_max)
        if self.world > 1:
            import torch.distributed as dist
            t = torch.as_tensor(sums, device=self.pool.device)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)  # MCNS.py:685
            sums = t.cpu().numpy()
        avg = sums[:6] / max(sums[6], 1.0)
        self.set_steps(update_steps(self.steps(), avg, self.pool.move_ratio, lower_bound, upper_bound))
        return avg

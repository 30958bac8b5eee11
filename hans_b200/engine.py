"""WalkerPool: device-resident walker state + thin calls into libhans_b200.so.

This is the host side of the hot path.  It owns the device buffers (torch CUDA tensors, used only
as memory) that replace hs_alkane's module arrays ``hmatrix(3,3,nbox)`` / ``Rchain(3,nbeads,
nchains,nbox)``, and hands their pointers to the C ABI.  Every compute
method enqueues CUDA work on the current torch stream and returns without synchronising unless
it has to hand numpy data back.

Reference: NesSa/MCNS.py (MC_run :276-392, initialise_sim_cells :582-606, create_initial_configs
:628-644) and mpihans.py:210-245 for the nested-sampling step.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import DBL_MAX, DECISION_DTYPE, MODE_FUSED, PROPOSAL_DTYPE, Cfg, check, lib, make_cfg


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class WalkerPool:
    """`nwalkers` walkers of `nchains` chains x `nbeads` beads resident on one GPU."""

    def __init__(self, nwalkers: int, nchains: int, nbeads: int, move_ratio: Sequence[float],
                 device: Optional[torch.device] = None, seed: int = 0, gid_base: int = 0,
                 threads_per_walker: int = 0, workspace_mb: int = 1024, **cfg_kw):
        if not torch.cuda.is_available():
            raise _lib.HansError("hans_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.nwalkers, self.nchains, self.nbeads = int(nwalkers), int(nchains), int(nbeads)
        self.N = self.nchains * self.nbeads
        self.seed, self.gid_base = int(seed), int(gid_base)
        self.tpw = int(threads_per_walker)
        self.move_ratio = np.asarray(move_ratio, dtype=np.float64)
        self._cfg_kw = dict(cfg_kw)
        self.cfg: Cfg = make_cfg(nchains, nbeads, self.move_ratio, **cfg_kw)
        f64 = dict(dtype=torch.float64, device=self.device)
        self.H = torch.zeros((self.nwalkers, 3, 3), **f64)
        self.R = torch.zeros((self.nwalkers, self.N, 3), **f64)
        self.ctr = torch.zeros(self.nwalkers, dtype=torch.int64, device=self.device)  # uint64 bits
        self.vol = torch.zeros(self.nwalkers, **f64)
        self.pair_tests = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.moves_per_sweep = _lib.moves_per_sweep(nbeads, nchains)
        self.launches = 0  # kernels launched through this pool (bench.py reports it)
        # proposals of a walk are generated ahead of it into this scratch (hans_mc_run_batch_ws);
        # 0 = generate inside the walk kernel
        self.workspace_bytes = int(workspace_mb) << 20
        self._ws: Optional[torch.Tensor] = None

    # ---- configuration ------------------------------------------------------------------------
    def set_move_ratio(self, move_ratio) -> None:
        self.move_ratio = np.asarray(move_ratio, dtype=np.float64)
        cum = np.cumsum(self.move_ratio) / np.sum(self.move_ratio)
        for k in range(6):
            self.cfg.move_cum[k] = float(cum[k])

    def cfg_with(self, move_ratio=None) -> Cfg:
        """A copy of the current cfg, optionally with another move ratio (adjust_mc_steps uses
        single-type ratios, MCNS.py:679-683)."""
        c = Cfg()
        C.memmove(C.byref(c), C.byref(self.cfg), C.sizeof(Cfg))
        if move_ratio is not None:
            mr = np.asarray(move_ratio, dtype=np.float64)
            cum = np.cumsum(mr) / np.sum(mr)
            for k in range(6):
                c.move_cum[k] = float(cum[k])
        return c

    # ---- state transfer -----------------------------------------------------------------------
    def set_state(self, H: np.ndarray, R: np.ndarray, ids: Optional[Sequence[int]] = None) -> None:
        H = torch.as_tensor(np.ascontiguousarray(H, dtype=np.float64)).reshape(-1, 3, 3)
        R = torch.as_tensor(np.ascontiguousarray(R, dtype=np.float64)).reshape(-1, self.N, 3)
        if ids is None:
            self.H.copy_(H.to(self.device))
            self.R.copy_(R.to(self.device))
        else:
            idx = torch.as_tensor(np.asarray(ids, dtype=np.int64), device=self.device)
            self.H[idx] = H.to(self.device)
            self.R[idx] = R.to(self.device)
        self.refresh_volumes()

    def get_state(self, ids: Optional[Sequence[int]] = None):
        if ids is None:
            return self.H.cpu().numpy(), self.R.cpu().numpy()
        idx = torch.as_tensor(np.asarray(ids, dtype=np.int64), device=self.device)
        return self.H[idx].cpu().numpy(), self.R[idx].cpu().numpy()

    def _ids(self, ids) -> Optional[torch.Tensor]:
        if ids is None:
            return None
        if isinstance(ids, torch.Tensor):
            return ids.to(device=self.device, dtype=torch.int32).contiguous()
        return torch.as_tensor(np.asarray(ids, dtype=np.int32), device=self.device)

    # ---- cell predicates ----------------------------------------------------------------------
    def refresh_volumes(self) -> torch.Tensor:
        """alk.box_compute_volume for every walker (mpihans.py:183)."""
        check(lib.hans_cell_metrics(self.H.data_ptr(), self.nwalkers, self.vol.data_ptr(), None, None, _stream()))
        self.launches += 1
        return self.vol

    def cell_metrics(self):
        """(volume, min_aspect_ratio, max |cos|) per walker (MCNS.py:97-133)."""
        f64 = dict(dtype=torch.float64, device=self.device)
        v, a, c = (torch.empty(self.nwalkers, **f64) for _ in range(3))
        check(lib.hans_cell_metrics(self.H.data_ptr(), self.nwalkers, v.data_ptr(), a.data_ptr(), c.data_ptr(), _stream()))
        self.launches += 1
        return v, a, c

    def check_overlap(self, ids=None, inter_only: bool = False) -> torch.Tensor:
        """alk.alkane_check_chain_overlap (MCNS.py:198,247,543) for `ids` (default: all)."""
        d_ids = self._ids(ids)
        n = self.nwalkers if d_ids is None else int(d_ids.numel())
        out = torch.empty(n, dtype=torch.int32, device=self.device)
        check(lib.hans_check_overlap(C.byref(self.cfg), self.H.data_ptr(), self.R.data_ptr(), _ptr(d_ids), n,
                                     int(inter_only), out.data_ptr(), self.tpw, _stream()))
        self.launches += 1
        return out

    def change_box(self, ids, dH: np.ndarray) -> None:
        """alk.alkane_change_box (MCNS.py:186,238,368)."""
        d_ids = self._ids(ids)
        n = int(d_ids.numel())
        d = torch.as_tensor(np.ascontiguousarray(dH, dtype=np.float64).reshape(n, 9)).to(self.device)
        check(lib.hans_change_box(C.byref(self.cfg), self.H.data_ptr(), self.R.data_ptr(), d_ids.data_ptr(), n,
                                  d.data_ptr(), _stream()))
        self.launches += 1

    def order_parameters(self):
        """(P2 nematic, translational) order parameter of every walker (nematic.py:30-137), device tensors."""
        f64 = dict(dtype=torch.float64, device=self.device)
        p2, tr = torch.empty(self.nwalkers, **f64), torch.empty(self.nwalkers, **f64)
        check(lib.hans_order_params(C.byref(self.cfg), self.H.data_ptr(), self.R.data_ptr(), self.nwalkers,
                                    p2.data_ptr(), tr.data_ptr(), _stream()))
        self.launches += 1
        return p2, tr

    # ---- initial configurations ---------------------------------------------------------------
    def set_initial_cells(self, max_vol_per_atom: float = 15.0) -> None:
        """create_initial_configs, MCNS.py:628-631."""
        cell = 0.999 * np.eye(3) * np.cbrt(self.nbeads * self.nchains * max_vol_per_atom)
        self.H.copy_(torch.as_tensor(cell).to(self.device).expand(self.nwalkers, 3, 3))

    def grow(self, max_attempts: int = 100000) -> np.ndarray:
        """populate_boxes (MCNS.py:634-644) for every walker; returns attempts used per walker."""
        status = torch.empty(self.nwalkers, dtype=torch.int32, device=self.device)
        check(lib.hans_grow(C.byref(self.cfg), self.H.data_ptr(), self.R.data_ptr(), self.nwalkers, self.seed,
                            self.gid_base, int(max_attempts), status.data_ptr(), _stream()))
        self.launches += 1
        st = status.cpu().numpy()
        if (st < 0).any():
            raise _lib.HansError("chain growth failed for walkers %s" % np.nonzero(st < 0)[0][:8])
        self.refresh_volumes()
        return st

    # ---- the hot path -------------------------------------------------------------------------
    def mc_run(self, sweeps: int, ids=None, volume_limit: float = DBL_MAX,
               d_volume_limit: Optional[torch.Tensor] = None, counters: bool = True, cfg: Optional[Cfg] = None,
               count_pairs: bool = True):
        """MC_run (MCNS.py:276-392) for walkers `ids` (default: all) in one launch.

        Returns (attempted, accepted) int64 device tensors [n,6] (or None, None); volumes of the
        walked walkers are refreshed in self.vol.  Asynchronous."""
        d_ids = self._ids(ids)
        n = self.nwalkers if d_ids is None else int(d_ids.numel())
        att = acc = None
        if counters:
            att = torch.zeros((n, 6), dtype=torch.int64, device=self.device)
            acc = torch.zeros((n, 6), dtype=torch.int64, device=self.device)
        c = self.cfg if cfg is None else cfg
        nmoves = int(sweeps) * self.moves_per_sweep
        need = n * nmoves * PROPOSAL_DTYPE.itemsize
        ws_bytes = min(need, self.workspace_bytes)
        if ws_bytes and (self._ws is None or self._ws.numel() < ws_bytes):
            self._ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        check(lib.hans_mc_run_batch_ws(C.byref(c), self.H.data_ptr(), self.R.data_ptr(), self.ctr.data_ptr(), _ptr(d_ids),
                                       n, int(sweeps), _ptr(d_volume_limit), float(volume_limit), self.seed, self.gid_base,
                                       _ptr(att), _ptr(acc), self.vol.data_ptr(),
                                       self.pair_tests.data_ptr() if count_pairs else None, self.tpw,
                                       _ptr(self._ws) if ws_bytes else None, int(ws_bytes), _stream()))
        self.launches += lib.hans_mc_run_launches(n, nmoves, 1 if ws_bytes else 0, int(ws_bytes))  # (generate + walk) per chunk
        return att, acc

    def replay(self, ids, proposals: np.ndarray, volume_limit: float = DBL_MAX, mode: int = MODE_FUSED) -> np.ndarray:
        """Apply explicit proposals [n, nmoves] to walkers `ids`; returns decisions [n, nmoves]."""
        d_ids = self._ids(ids)
        n = self.nwalkers if d_ids is None else int(d_ids.numel())
        proposals = np.ascontiguousarray(proposals)
        assert proposals.dtype == PROPOSAL_DTYPE
        proposals = proposals.reshape(n, -1)
        nmoves = proposals.shape[1]
        d_p = torch.as_tensor(proposals.view(np.uint8).reshape(-1)).to(self.device)
        d_o = torch.zeros(n * nmoves * DECISION_DTYPE.itemsize, dtype=torch.uint8, device=self.device)
        check(lib.hans_replay(C.byref(self.cfg), self.H.data_ptr(), self.R.data_ptr(), _ptr(d_ids), n, d_p.data_ptr(),
                              nmoves, float(volume_limit), int(mode), d_o.data_ptr(), self.tpw, _stream()))
        self.launches += 1
        return d_o.cpu().numpy().view(DECISION_DTYPE).reshape(n, nmoves)

    # ---- nested-sampling helpers (mpihans.py:210-245) -------------------------------------------
    def argmax(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Local MAXLOC over the pool's volumes -> device tensor [max volume, index] (f64)."""
        if out is None:
            out = torch.empty(2, dtype=torch.float64, device=self.device)
        check(lib.hans_argmax(self.vol.data_ptr(), self.nwalkers, out.data_ptr(), _stream()))
        self.launches += 1
        return out

    def clone(self, src: int, dst: int, src_pool: Optional["WalkerPool"] = None,
              d_src: Optional[torch.Tensor] = None, d_dst: Optional[torch.Tensor] = None,
              d_enable: Optional[torch.Tensor] = None, srcH=None, srcR=None, srcvol=None) -> None:
        """Copy walker `src` (of src_pool or of raw staging tensors) over walker `dst` of this pool."""
        sp = self if src_pool is None else src_pool
        sH = sp.H if srcH is None else srcH
        sR = sp.R if srcR is None else srcR
        sv = sp.vol if (srcvol is None and srcH is None) else srcvol
        check(lib.hans_clone(C.byref(self.cfg), sH.data_ptr(), sR.data_ptr(), _ptr(sv), int(src), _ptr(d_src),
                             self.H.data_ptr(), self.R.data_ptr(), self.vol.data_ptr(), int(dst), _ptr(d_dst),
                             _ptr(d_enable), _stream()))
        self.launches += 1


def ffma_peak_tflops(iters: int = 10000) -> float:
    """Measured FP32 FFMA issue peak of the current GPU (roofline denominator in bench.py)."""
    out = C.c_double(0.0)
    check(lib.hans_ffma_peak(int(iters), C.byref(out), _stream()))
    return out.value

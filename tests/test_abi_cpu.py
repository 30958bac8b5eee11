"""The C-ABI library must load without a GPU and export every symbol include/hans_b200.h declares;
struct layouts on the Python side must match the header; argument validation must not need a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hans_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hans_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from hans_b200 import _lib
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in include/hans_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names, "hans_b200/_lib.py EXPORTS out of sync with the header"
    assert _lib.lib.hans_abi_version() == 1
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for n in names:
        assert re.search(rf"\bT {n}\b", out), n


def test_struct_layouts_match_header(tmp_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof with ctypes."""
    from hans_b200 import _lib
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "hans_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu ", sizeof(hans_cfg), sizeof(hans_proposal), sizeof(hans_decision), sizeof(hans_ns_args));
  printf("%zu %zu %zu %zu ", offsetof(hans_cfg, dv_max), offsetof(hans_cfg, move_cum), offsetof(hans_cfg, bondlength), offsetof(hans_cfg, sin_bondangle));
  printf("%zu %zu %zu ", offsetof(hans_proposal, p), offsetof(hans_proposal, u_acc), offsetof(hans_decision, boltz));
  printf("%zu %zu %zu\n", offsetof(hans_ns_args, seed), offsetof(hans_ns_args, d_H), offsetof(hans_ns_args, d_traj_iter));
  return 0; }''')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(_lib.Cfg), _lib.PROPOSAL_DTYPE.itemsize, _lib.DECISION_DTYPE.itemsize, C.sizeof(_lib.NsArgs),
            _lib.Cfg.dv_max.offset, _lib.Cfg.move_cum.offset, _lib.Cfg.bondlength.offset, _lib.Cfg.sin_bondangle.offset,
            _lib.PROPOSAL_DTYPE.fields["p"][1], _lib.PROPOSAL_DTYPE.fields["u_acc"][1], _lib.DECISION_DTYPE.fields["boltz"][1],
            _lib.NsArgs.seed.offset, _lib.NsArgs.d_H.offset, _lib.NsArgs.d_traj_iter.offset]
    assert got == want
    # the oracle's structs are laid out identically (same proposals feed both)
    from oracle import oracle as ho
    assert C.sizeof(ho.Cfg) == C.sizeof(_lib.Cfg) and ho.PROPOSAL_DTYPE == _lib.PROPOSAL_DTYPE


def test_host_side_entry_points_and_validation():
    from hans_b200 import _lib
    L = _lib.lib
    assert L.hans_moves_per_sweep(6, 16) == 87 and L.hans_moves_per_sweep(1, 64) == 71    # MCNS.py:296-301
    cfg = _lib.make_cfg(16, 6, [3, 48, 32, 0, 3, 3])
    assert list(cfg.move_cum)[-1] == 1.0 and abs(cfg.move_cum[0] - 3 / 89) < 1e-15
    # f64 master + trial + cell (8 (6N + 3nb + 36)), float4 shadows + filter (16 (2N + nb) + 128), chain radii
    # (4 nchains), near lists (16 + 8 T), ring (64 T)
    want = ((8 * (6 * 96 + 18 + 36) + 15) // 16 * 16 + 16 * (2 * 96 + 6) + 128 + 4 * 16 + 16 + 8 * 32 + 63) // 64 * 64 + 64 * 32
    assert L.hans_walker_smem_bytes(C.byref(cfg)) == want
    assert L.hans_default_threads_per_walker(C.byref(cfg)) == 32
    big = _lib.make_cfg(128, 6, [1, 384, 256, 384, 3, 3])
    assert L.hans_default_threads_per_walker(C.byref(big)) == 128
    # kernel selection per shape: warp-per-walker kernel for small walkers and for chain systems
    # of up to 64 chains, the cooperative one otherwise
    assert L.hans_default_window(C.byref(cfg)) == 1
    assert L.hans_default_window(C.byref(_lib.make_cfg(64, 1, [1, 192, 0, 0, 3, 3]))) == 1
    assert L.hans_default_window(C.byref(_lib.make_cfg(32, 12, [3, 32, 32, 16, 3, 3]))) == 0   # 384 beads: CTA of 64
    assert L.hans_default_threads_per_walker(C.byref(_lib.make_cfg(32, 12, [3, 32, 32, 16, 3, 3]))) == 64
    assert L.hans_default_window(C.byref(_lib.make_cfg(16, 12, [3, 32, 32, 16, 3, 3]))) == 1
    assert L.hans_default_window(C.byref(big)) == 0                                        # 128 chains: one CTA per walker
    assert L.hans_default_team(C.byref(big)) == 8                                          # ... of 8 warps, one move each
    assert L.hans_default_team(C.byref(cfg)) == 0 and L.hans_default_team(C.byref(_lib.make_cfg(32, 12, [3, 32, 32, 16, 3, 3]))) == 4
    assert L.hans_default_window(C.byref(_lib.make_cfg(16, 5, [1, 1, 1, 1, 1, 1]))) == 0   # nbeads = 5 is not instantiated
    assert L.hans_default_window(C.byref(_lib.make_cfg(200, 1, [1, 1, 0, 0, 1, 1]))) == 0  # 200 spheres
    assert L.hans_mc_run_batch_ws(None, None, None, None, None, 1, 1, None, 0.0, 0, 0, None, None, None, None, 0, None, 0, None) == -1
    assert L.hans_gen_proposals(C.byref(cfg), None, None, 1, 1, 0, 0, None, None) == -1
    assert b"required" in L.hans_last_error()
    # argument errors are reported as codes + text, never by exiting, and need no GPU
    assert L.hans_mc_run_batch(None, None, None, None, None, 1, 1, None, 0.0, 0, 0, None, None, None, None, 0, None) == -1
    assert b"cfg" in L.hans_last_error()
    bad = _lib.make_cfg(4, 100, [1, 1, 1, 1, 1, 1])
    assert L.hans_check_overlap(C.byref(bad), 1, 1, None, 1, 0, 1, 0, None) == -3          # HANS_ELIMIT: nbeads > 64
    assert L.hans_replay(C.byref(cfg), 1, 1, None, 1, 1, 1, 0.0, 7, 1, 0, None) == -1       # bad mode
    assert L.hans_mc_run_batch(C.byref(cfg), 1, 1, 1, None, 1, 1, None, 0.0, 0, 0, None, None, None, None, 48, None) == -1
    with pytest.raises(_lib.HansError):
        _lib.check(-1)
    with pytest.raises(Exception):
        _lib.make_cfg(4, 2, [1, 2, 3])                                                      # MCNS.py:614-615
    assert L.hans_device_count() >= 0


def test_empty_inputs_are_no_ops_and_negative_counts_are_errors():
    """Zero walkers / zero moves: every batched entry point returns HANS_OK before it touches the device (so this runs
    without a GPU; the pointers are never dereferenced); a negative count is HANS_EINVAL."""
    from hans_b200 import _lib
    L = _lib.lib
    cfg = _lib.make_cfg(16, 6, [3, 48, 32, 0, 3, 3])
    c, p = C.byref(cfg), C.c_void_p(4096)   # any non-NULL "device pointer"
    assert L.hans_mc_run_batch_ws(c, p, p, p, None, 0, 40, None, 1e300, 1, 0, None, None, None, None, 0, None, 0, None) == 0
    assert L.hans_mc_run_batch(c, p, p, p, None, 0, 40, None, 1e300, 1, 0, None, None, None, None, 0, None) == 0
    assert L.hans_gen_proposals(c, p, None, 0, 87, 1, 0, p, None) == 0 and L.hans_gen_proposals(c, p, None, 5, 0, 1, 0, p, None) == 0
    assert L.hans_replay(c, p, p, None, 0, p, 10, 1e300, 0, p, 0, None) == 0 and L.hans_replay(c, p, p, None, 3, p, 0, 1e300, 0, p, 0, None) == 0
    assert L.hans_check_overlap(c, p, p, None, 0, 0, p, 0, None) == 0
    assert L.hans_cell_metrics(p, 0, p, p, p, None) == 0
    assert L.hans_change_box(c, p, p, None, 0, p, None) == 0
    assert L.hans_shape_delta(p, None, 0, p, p, None) == 0
    assert L.hans_grow(c, p, p, 0, 1, 0, 10, None, None) == 0
    assert L.hans_order_params(c, p, p, 0, p, p, None) == 0
    assert L.hans_mc_run_launches(0, 3480, 1, 1 << 30) == 0
    for rc in (L.hans_mc_run_batch_ws(c, p, p, p, None, -1, 40, None, 1e300, 1, 0, None, None, None, None, 0, None, 0, None),
               L.hans_mc_run_batch_ws(c, p, p, p, None, 4, -1, None, 1e300, 1, 0, None, None, None, None, 0, None, 0, None),
               L.hans_gen_proposals(c, p, None, -1, 1, 1, 0, p, None), L.hans_replay(c, p, p, None, -2, p, 1, 1e300, 0, p, 0, None),
               L.hans_check_overlap(c, p, p, None, -1, 0, p, 0, None), L.hans_cell_metrics(p, -1, p, p, p, None),
               L.hans_change_box(c, p, p, None, -1, p, None), L.hans_shape_delta(p, None, -1, p, p, None),
               L.hans_grow(c, p, p, -1, 1, 0, 10, None, None), L.hans_grow(c, p, p, 0, 1, 0, 0, None, None),
               L.hans_order_params(c, p, p, -1, p, p, None)):
        assert rc == -1, L.hans_last_error()
    # attempted / accepted counters go together (checked before the early return)
    assert L.hans_mc_run_batch_ws(c, p, p, p, None, 0, 40, None, 1e300, 1, 0, p, None, None, None, 0, None, 0, None) == -1


def test_no_product_module_touches_the_oracle():
    """The product path must not import, link, open or execute anything under oracle/ (nor the test harness)."""
    import ast
    pkg = os.path.join(ROOT, "hans_b200")
    banned_modules = ("oracle", "tests", "ref_live", "common")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith((".cu", ".cuh", ".h")):
                text = open(path).read()
                for m in re.finditer(r'#\s*include\s*[<"]([^>"]+)[>"]', text):
                    assert "oracle" not in m.group(1).lower(), f"{f} includes {m.group(1)}"
                assert "libhans_oracle" not in text and not re.search(r"\bho_\w+\s*\(", text), f"{f} calls into the checker"
                continue
            if not f.endswith(".py"):
                continue
            text = open(path).read()
            tree = ast.parse(text)
            for node in ast.walk(tree):
                if isinstance(node, ast.Import):
                    for a in node.names:
                        assert a.name.split(".")[0] not in banned_modules, f"{f} imports {a.name}"
                elif isinstance(node, ast.ImportFrom):
                    mod = node.module or ""
                    assert mod.split(".")[0] not in banned_modules or node.level > 0, f"{f} imports from {mod}"
                    if node.level > 0:   # relative import inside hans_b200: must not climb out of the package
                        assert node.level == 1, f"{f} uses a relative import that leaves the package"
                elif isinstance(node, ast.Call):
                    # dynamic imports / library loads: every string literal argument is checked
                    fn = node.func
                    name = fn.attr if isinstance(fn, ast.Attribute) else getattr(fn, "id", "")
                    if name in ("__import__", "import_module", "CDLL", "LoadLibrary", "spec_from_file_location", "run", "Popen",
                                "system", "exec", "eval"):
                        for a in ast.walk(node):
                            if isinstance(a, ast.Constant) and isinstance(a.value, str):
                                assert "oracle" not in a.value.lower(), f"{f}: {name}(... {a.value!r} ...)"
                        assert name not in ("exec", "eval"), f"{f} uses {name}()"
            # and no string anywhere that names the checker's files
            for node in ast.walk(tree):
                if isinstance(node, ast.Constant) and isinstance(node.value, str) and node.value != ast.get_docstring(tree):
                    low = node.value.lower()
                    assert "libhans_oracle" not in low and "oracle/" not in low and "hans_oracle" not in low, f"{f}: {node.value[:60]!r}"
    # the only library the package loads is its own
    from hans_b200 import _lib
    assert os.path.basename(_lib.LIB_PATH) == "libhans_b200.so" or "HANS_B200_LIB" in os.environ


def test_workspace_plan_launch_count():
    """hans_mc_run_launches mirrors how hans_mc_run_batch_ws cuts a walk (include/hans_b200.h)."""
    from hans_b200 import _lib
    L = _lib.lib
    per_move = 1024 * 64
    assert L.hans_mc_run_launches(1024, 2840, 1, 2840 * per_move) == 2        # whole walk fits: generate, walk
    assert L.hans_mc_run_launches(1024, 2840, 1, 10 ** 12) == 2
    assert L.hans_mc_run_launches(1024, 2840, 1, 1000 * per_move) == 2 * 3    # 992 moves fit: generate, walk, three times
    assert L.hans_mc_run_launches(1024, 2840, 0, 0) == 1                       # no workspace: generation inside the walk kernel
    assert L.hans_mc_run_launches(8, 87, 1, 8 * 64 * 40) == 2 * 3              # 32 of 87 moves fit


def test_peer_exchange_buffer_size():
    """hans_ns_peer_bytes: flags[2][world] u64 padded to 128 bytes + records[2][world][12 + 3N] float64."""
    from hans_b200 import _lib
    L = _lib.lib
    assert L.hans_ns_peer_bytes(96, 8) == 8 * (16 + 2 * 8 * (12 + 288))
    assert L.hans_ns_peer_bytes(768, 2) == 8 * (16 + 2 * 2 * (12 + 2304))
    assert L.hans_ns_peer_bytes(64, 16) == 8 * (32 + 2 * 16 * (12 + 192))
    assert L.hans_ns_peer_bytes(0, 2) == 0


def test_product_fails_loudly_without_extension_or_gpu():
    """No CPU fallback anywhere on the product path: a missing library is an ImportError that says so; without a CUDA
    device the walker pool refuses to exist; the GPU arm of bench.py exits non-zero instead of measuring something else."""
    import subprocess
    import sys
    env = dict(os.environ, HANS_B200_LIB=os.path.join(ROOT, "does", "not", "exist.so"))
    r = subprocess.run([sys.executable, "-c", "import hans_b200._lib"], capture_output=True, text=True, cwd=ROOT, env=env, timeout=300)
    assert r.returncode != 0 and "ImportError" in r.stderr and "no CPU fallback" in r.stderr
    import torch
    if not torch.cuda.is_available():
        from hans_b200 import _lib, engine
        with pytest.raises(_lib.HansError, match="no CPU fallback"):
            engine.WalkerPool(4, 16, 6, [3, 48, 32, 0, 3, 3])
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--no-cpu"],
                           capture_output=True, text=True, cwd=ROOT, timeout=600)
        assert r.returncode != 0 and not r.stdout.strip().startswith("{"), (r.returncode, r.stdout[:200])

"""CPU tests of the product's host logic and of the integer code shared with the kernels (tests/harness):
index maps, O(1) fingerprint composition, band tables, the register-DP row step, score arithmetic, TSV I/O,
and that libnwsolve.so loads and exports every symbol include/nwsolve.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from nahrwhals_b200.sdgrid import random_grid, sdgrid, to_solver_inputs, write_tsvs
from oracle import nw_oracle_py as op

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def H():
    d = os.path.join(ROOT, "tests", "harness")
    subprocess.check_call(["make", "-s", "-C", d])
    return C.CDLL(os.path.join(d, "libnw_harness.so"))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _child(H, par, kind, p, q):
    par = np.ascontiguousarray(par, np.int16)
    out = np.zeros(2 * len(par) + 2, np.int16)
    n = H.h_child(_p(par), len(par), kind, p, q, _p(out))
    return out[:n].tolist()


def _fast(H, aln, lx, ly, par, kind, p, q, force):
    par = np.ascontiguousarray(par, np.int16)
    cost, tot, win = C.c_int64(), C.c_int64(), C.c_int()
    rc = H.h_score_fast(_p(aln), aln.shape[0], aln.shape[1], _p(lx), _p(ly), _p(par), len(par), kind, p, q, force,
                        C.byref(cost), C.byref(tot), C.byref(win))
    return rc, cost.value, tot.value, win.value


def test_index_map_matches_apply_functions(H):
    """child_elem/child_len vs apply_duplication/deletion/inversion (solver.jl:244-278)."""
    rng = np.random.default_rng(5)
    for _ in range(300):
        L = int(rng.integers(2, 40))
        par = tuple(int(v) for v in rng.integers(1, 30, L) * rng.choice([-1, 1], L))
        p = int(rng.integers(1, L))
        q = int(rng.integers(p + 1, L + 1))
        for kind in (0, 1, 2):
            assert tuple(_child(H, par, kind, p, q)) == op.apply_move(kind, par, p, q)
        assert tuple(_child(H, par, 3, p, q)) == par


def test_fingerprint_composition(H):
    """O(1) child fingerprint from the parent's prefix tables == direct fingerprint of the materialised child,
    and distinct children of one parent get distinct fingerprints."""
    rng = np.random.default_rng(6)
    for _ in range(60):
        L = int(rng.integers(2, 120))
        par = np.ascontiguousarray(rng.integers(1, 50, L) * rng.choice([-1, 1], L), np.int16)
        seen = {}
        for _ in range(40):
            p = int(rng.integers(1, L))
            q = int(rng.integers(p + 1, L + 1))
            kind = int(rng.integers(0, 4))
            assert H.h_fp_check(_p(par), L, kind, p, q) == 1
            ch = tuple(_child(H, par, kind, p, q))
            fp = np.zeros(4, np.uint32)
            H.h_fp(_p(np.ascontiguousarray(ch, np.int16)), len(ch), _p(fp))
            key = tuple(fp.tolist())
            assert seen.setdefault(key, ch) == ch  # equal fingerprint => equal sequence
            assert all(v < 2**31 - 1 for v in key)


def test_band_cells_match_oracle(H, oracle):
    """cells(L, n_y) of nw_band.h == number of DP cells the oracle evaluates (SURVEY 8d definition)."""
    for n_y in (1, 2, 5, 9, 10, 11, 12, 30, 64, 97):
        aln = np.ones((3, n_y), np.int8)
        lx = np.ones(3)
        ly = np.ones(n_y)
        Ls = [L for L in range(1, 3 * n_y + 12)]
        seqs = [[1 + (t % 3) for t in range(L)] for L in Ls]
        for force in (0, 1):
            _, cost, _, cells = oracle.score_batch(aln, lx, ly, seqs, force=bool(force))
            for L, c, cc in zip(Ls, cells, cost):
                out = C.c_int64()
                a, b, d = C.c_int(), C.c_int(), C.c_int()
                H.h_band(L, n_y, force, C.byref(out), C.byref(a), C.byref(b), C.byref(d))
                assert d.value == 1
                if np.isnan(cc):
                    continue  # early exit: no DP
                assert out.value == c, (L, n_y, force)


def test_band_rows_pointer_construction(H):
    """O(L + n_y) construction of the corridor rows == exhaustive evaluation of every cell with the Float64
    predicates of solver.jl:390/:400/:412-413, for every shape class incl. the reference's 250-gridline cap."""
    for n_y in list(range(1, 70)) + [97, 128, 150, 200, 249, 250, 333]:
        for L in range(1, 3 * n_y + 12, 1 if n_y < 40 else 5):
            for force in (0, 1):
                assert H.h_band_check(L, n_y, force) == 1, (L, n_y, force)


def test_register_dp_matches_oracle(H, oracle):
    """dp_row_step (the code the sm_100a kernel runs per row) vs the oracle's Float64 DP, bit-exact cost/total."""
    rng = np.random.default_rng(0)
    nfast = 0
    windows = set()
    for trial in range(250):
        n_x = int(rng.integers(3, 90))
        n_y = int(rng.integers(1, 90))
        if trial % 3 == 0:
            m, lx, ly = random_grid(n_x, n_y, 0.08, trial)
        else:
            fams = [int(rng.integers(2, 6)) for _ in range(int(rng.integers(1, 5)))]
            if sum(fams) > n_x:
                continue
            m, lx, ly, _ = sdgrid(n_x, fams, seed=trial, chain_depth=int(rng.integers(0, 4)))
        aln = to_solver_inputs(m)
        n_x = aln.shape[0]
        par = list(range(1, n_x + 1))
        for _ in range(int(rng.integers(0, 3))):
            L = len(par)
            if L < 3:
                break
            a = int(rng.integers(1, L))
            b = int(rng.integers(a + 1, L + 1))
            k = int(rng.integers(0, 3))
            if k == 0 and L + (b - a) > 3 * n_x:
                continue
            par = _child(H, par, k, a, b)
        descs, seqs = [], []
        for _ in range(6):
            L = len(par)
            if L < 2:
                break
            a = int(rng.integers(1, L))
            b = int(rng.integers(a + 1, L + 1))
            k = int(rng.integers(0, 4))
            descs.append((k, a, b))
            seqs.append(_child(H, par, k, a, b))
        for force in (0, 1):
            sc, co, to, _ = oracle.score_batch(aln, lx, ly, seqs, force=bool(force))
            for (k, a, b), c0, t0, s0 in zip(descs, co, to, sc):
                rc, cost, tot, win = _fast(H, aln, lx, ly, par, k, a, b, force)
                if rc == -1:
                    assert np.isnan(c0)
                elif rc == 1:
                    nfast += 1
                    windows.add(win)
                    assert (cost, tot) == (c0, t0)
                    assert H.h_score_k3(int(cost), int(tot)) == int(round(s0 * 1000))
    assert nfast > 1000 and len(windows) >= 8


def test_packed_register_dp_matches_oracle(H, oracle):
    """dp_row_step2 (two candidates of one child length in the 16-bit halves of a register, the code k_score_pair runs
    per row) vs the oracle's Float64 DP, bit-exact cost/total for both halves; totals close to the 15-bit limit
    included, and shapes beyond it rejected by pair_range_ok."""
    rng = np.random.default_rng(11)
    npair = 0
    nreject = 0
    windows = set()
    for trial in range(260):
        n_x = int(rng.integers(3, 90))
        n_y = int(rng.integers(1, 90))
        if trial % 3 == 0:
            m, lx, ly = random_grid(n_x, n_y, 0.08, trial)
        else:
            fams = [int(rng.integers(2, 6)) for _ in range(int(rng.integers(1, 5)))]
            if sum(fams) > n_x:
                continue
            m, lx, ly, _ = sdgrid(n_x, fams, seed=trial, chain_depth=int(rng.integers(0, 4)))
        if trial % 5 == 4:  # scale the units so that totals approach (and sometimes exceed) the 15-bit range
            f = int(rng.integers(2, 30))
            lx, ly = lx * f, ly * f
            lx[0] += 1000 * int(lx[0] // 1000 == lx[0] / 1000)  # keep the gcd small
        aln = to_solver_inputs(m)
        n_x = aln.shape[0]
        par = list(range(1, n_x + 1))
        for _ in range(int(rng.integers(0, 3))):
            L = len(par)
            if L < 3:
                break
            a = int(rng.integers(1, L))
            b = int(rng.integers(a + 1, L + 1))
            k = int(rng.integers(0, 3))
            if k == 0 and L + (b - a) > 3 * n_x:
                continue
            par = _child(H, par, k, a, b)
        L = len(par)
        if L < 4:
            continue
        parr = np.ascontiguousarray(par, np.int16)
        for _ in range(8):
            # two moves with the same child length: same kind and same q - p for dup / del, anything for inv / id
            k = int(rng.integers(0, 4))
            if k in (0, 1):
                d = int(rng.integers(1, L))
                pa = int(rng.integers(1, L - d + 1))
                pb = int(rng.integers(1, L - d + 1))
                moves = [(k, pa, pa + d), (k, pb, pb + d)]
            else:
                moves = []
                for _ in range(2):
                    a = int(rng.integers(1, L))
                    moves.append((int(rng.choice([2, 3])), a, int(rng.integers(a + 1, L + 1))))
            seqs = [_child(H, par, *mv) for mv in moves]
            kind = np.array([mv[0] for mv in moves], np.int32)
            pp = np.array([mv[1] for mv in moves], np.int32)
            qq = np.array([mv[2] for mv in moves], np.int32)
            for force in (0, 1):
                cost = np.zeros(2, np.int64)
                tot = np.zeros(2, np.int64)
                rc = H.h_score_pair(_p(aln), aln.shape[0], aln.shape[1], _p(lx), _p(ly), _p(parr), L, _p(kind), _p(pp),
                                    _p(qq), force, _p(cost), _p(tot))
                if rc != 1:
                    nreject += rc == 0
                    continue
                _, co, to, _ = oracle.score_batch(aln, lx, ly, seqs, force=bool(force))
                npair += 1
                assert cost.tolist() == [int(v) for v in co], (trial, moves, force)
                assert tot.tolist() == [int(v) for v in to]
    assert npair > 1500 and nreject > 0
    # totals right below the 15-bit limit: block lengths in bp with gcd 1, scaled so that sum_rt + L * max_up lands in
    # (31 000, 32 767] for the children scored
    nedge = 0
    for trial in range(40):
        n_x = int(rng.integers(12, 40))
        m, lx, ly = random_grid(n_x, n_x, 0.12, 500 + trial, len_choices=tuple(int(v) for v in rng.integers(1, 60, 6)) + (1,))
        lx, ly = lx // 1000, ly // 1000  # random_grid hands out multiples of 1000: back to small integers
        aln = to_solver_inputs(m)
        par = list(range(1, aln.shape[0] + 1))
        L = len(par)
        f = 32767 // (int(ly.sum()) + L * int(lx.max()))
        if f < 1:
            continue
        lx2, ly2 = lx * f, ly * f
        lx2[0] += 1 if (int(ly2.sum()) + L * int(max(lx2.max(), lx2[0] + 1)) <= 32767) else 0
        parr = np.ascontiguousarray(par, np.int16)
        for _ in range(6):
            moves = []
            for _ in range(2):
                a = int(rng.integers(1, L))
                moves.append((2, a, int(rng.integers(a + 1, L + 1))))
            seqs = [_child(H, par, *mv) for mv in moves]
            kind = np.array([mv[0] for mv in moves], np.int32)
            pp = np.array([mv[1] for mv in moves], np.int32)
            qq = np.array([mv[2] for mv in moves], np.int32)
            cost = np.zeros(2, np.int64)
            tot = np.zeros(2, np.int64)
            rc = H.h_score_pair(_p(aln), aln.shape[0], aln.shape[1], _p(lx2), _p(ly2), _p(parr), L, _p(kind), _p(pp), _p(qq),
                                1, _p(cost), _p(tot))
            if rc != 1:
                continue
            _, co, to, _ = oracle.score_batch(aln, lx2, ly2, seqs, force=True)
            assert cost.tolist() == [int(v) for v in co] and tot.tolist() == [int(v) for v in to]
            g = int(np.gcd.reduce(np.concatenate([lx2, ly2])))
            nedge += int(max(to)) // g > 20000
    assert nedge > 20
    # the range argument behind pair_range_ok: no 16-bit addition of the packed step ever wrapped, masked cells included
    H.h_pair_wraps.restype = C.c_long
    assert H.h_pair_wraps() == 0


def test_window_class_folding(H):
    """fold_window_classes (run_scoring's launch grouping): fold iff n * (B - A) / A < half a wave; folds chain with the
    accumulated work; the packed and the 32-bit kernel never mix; the generic class (0) is never folded."""
    PAIR = 0x1000
    wave = 148 * 6

    def fold(classes):
        cls = np.array([c for c, _ in classes], np.int32)
        slots = np.array([s for _, s in classes], np.uint64)
        out = np.zeros(len(classes), np.int32)
        H.h_fold_classes(len(classes), _p(cls), _p(slots), C.c_uint64(wave), _p(out))
        return out.tolist()

    # S150: window 60 (a few thousand CTAs) folds into 64: 4 / 60 more work each
    assert fold([(PAIR | 60, 128 * 3000), (PAIR | 64, 128 * 20000)]) == [PAIR | 64, PAIR | 64]
    # ... but not when it is a launch of many waves on its own
    assert fold([(PAIR | 60, 128 * 7000), (PAIR | 64, 128 * 20000)]) == [PAIR | 60, PAIR | 64]
    # a narrow class is kept out of a much wider window unless it is tiny (28 -> 128: 100 / 28 more work each)
    assert fold([(PAIR | 28, 128 * 200), (PAIR | 128, 128 * 50)]) == [PAIR | 28, PAIR | 128]
    assert fold([(PAIR | 28, 128 * 100), (PAIR | 128, 128 * 50)]) == [PAIR | 128, PAIR | 128]
    # chains accumulate: 40 -> 44, then 44 (with 40's work) is too big for 48
    n40, n44 = 128 * 3000, 128 * 3000
    assert fold([(PAIR | 40, n40), (PAIR | 44, n44), (PAIR | 48, 128 * 9000)]) == [PAIR | 44, PAIR | 44, PAIR | 48]
    assert fold([(PAIR | 40, 128 * 500), (PAIR | 44, 128 * 500), (PAIR | 48, 128 * 9000)]) == [PAIR | 48] * 3
    # kinds never mix, generic stays
    assert fold([(0, 128), (64, 128), (PAIR | 8, 128), (PAIR | 12, 128)]) == [0, 64, PAIR | 12, PAIR | 12]
    assert fold([(128, 128 * 10), (PAIR | 8, 128 * 10)]) == [128, PAIR | 8]


def test_score_string_all_values(H, oracle):
    """The product's TSV score formatter vs a generic shortest-round-trip Float32 printer, all 100001 values."""
    buf = C.create_string_buffer(64)
    for k3 in range(0, 100001, 7):
        H.h_score_string(k3, buf, 64)
        assert buf.value.decode() == oracle.f32_string(np.float32(k3 / 1000.0)), k3
    for k3 in (0, 1, 10, 100, 1000, 99595, 100000, 71378):
        H.h_score_string(k3, buf, 64)
        assert buf.value.decode() == oracle.f32_string(np.float32(k3 / 1000.0))


def test_library_exports_declared_symbols():
    """libnwsolve.so loads without a GPU and exports exactly the functions include/nwsolve.h declares."""
    from nahrwhals_b200 import _lib
    L = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "nwsolve.h")).read()
    declared = set(re.findall(r"\b(nw_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS)
    for s in declared:
        assert getattr(L, s) is not None
    assert b"sm_100a" in L.nw_version()
    from nahrwhals_b200 import grid, paf, region
    hdr = open(os.path.join(ROOT, "include", "nwpaf.h")).read()
    declared = set(re.findall(r"\b(nw_paf_[a-z_0-9]+)\s*\(", hdr)) - {"nw_paf_cols", "nw_paf_table"}
    assert declared == set(paf.SYMBOLS)
    for s in declared:
        assert getattr(L, s) is not None
    hdr = open(os.path.join(ROOT, "include", "nwgrid.h")).read()
    declared = set(re.findall(r"\b(nw_grid_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(grid.SYMBOLS)
    for s in declared:
        assert getattr(L, s) is not None
    from nahrwhals_b200 import bitlocus
    hdr = open(os.path.join(ROOT, "include", "nwbitlocus.h")).read()
    declared = set(re.findall(r"\b(nw_bitlocus_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(bitlocus.SYMBOLS)
    for s in declared:
        assert getattr(L, s) is not None
    from nahrwhals_b200 import pipeline
    hdr = open(os.path.join(ROOT, "include", "nwpipeline.h")).read()
    declared = set(re.findall(r"\b(nw_regions_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(pipeline.SYMBOLS)
    for s in declared:
        assert getattr(L, s) is not None
    hdr = open(os.path.join(ROOT, "include", "nwregion.h")).read()
    declared = set(re.findall(r"\b(nw_(?:region|table)_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(region.SYMBOLS)
    for s in declared:
        assert getattr(L, s) is not None


def test_read_problem_accepts_r_number_formats(tmp_path):
    """read_tsv / read_length_tsv (solver.jl:101-138): values like 1e+05 (R without scipen), two-line lens file."""
    import nahrwhals_b200 as nb
    m = tmp_path / "m.tsv"
    l = tmp_path / "l.tsv"
    m.write_text("1e+05\t0\t-30000\n0\t2e+04\t0\n")
    l.write_text("1e+05\t20000\n1e+05\t2e+04\t30000\n")
    aln, lx, ly = nb.read_problem(m, l)
    assert aln.tolist() == [[1, 0], [0, 1], [-1, 0]]
    assert lx.tolist() == [100000, 20000, 30000] and ly.tolist() == [100000, 20000]
    l.write_text("1\t2\n1\t2\t3\n4\n")
    with pytest.raises(Exception, match="exactly two rows"):
        nb.read_problem(m, l)
    m.write_text("1\t2\n3\n")
    with pytest.raises(Exception):
        nb.read_problem(m, l)


def test_no_gpu_fails_loudly():
    """Without a CUDA device the product must raise, never fall back to a CPU path."""
    import torch
    import nahrwhals_b200 as nb
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(Exception, match="CUDA"):
        nb.Solver(0)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the arm the driver runs next to ours): one JSON line with the contract's keys, on a
    small sample so that the CPU suite stays short."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup",
                          "0", "--cpu-sample", "3000"], capture_output=True, text=True, timeout=300, check=True).stdout
    assert len([l for l in out.split("\n") if l.strip()]) == 1  # stdout is exactly the JSON line
    line = json.loads(out.strip().split("\n")[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in line["config"]

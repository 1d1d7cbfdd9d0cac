"""save_to_logfile's row of nahrwhals_res.tsv (R/save_logfile.R:112-232): the C ABI (host side) against the restatement."""
import os

import numpy as np
import pytest

from oracle import nw_region_oracle as ro
from tests.test_region_oracle import load_testrun, oracle_solver  # noqa: F401

META = dict(chr="chr1_partial", start="1700000", end="3300000", samplename="Fasta_x_Fasta_y", xpad="1")


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle_lib
    return oracle_lib


def test_r_fixed15():
    for v, want in [(71.378, "71.378"), (99.595, "99.595"), (100.0, "100"), (1600000.0, "1600000"), (2e6, "2000000"),
                    (1e5, "100000"), (0.0, "0"), (float("nan"), "NA"), (1 / 3, "0.333333333333333"), (2810000.0, "2810000"),
                    (0.5, "0.5"), (1e-5, "0.00001"), (123456.789, "123456.789")]:
        assert ro.r_fixed15(v) == want, v


@pytest.mark.parametrize("seed", range(5))
def test_logrow_matches_oracle(oracle, seed, tmp_path):
    """format_julia_output tables (host only) -> summary -> log row, product vs restatement; header; append under lock."""
    import nahrwhals_b200 as nb
    from nahrwhals_b200 import region
    from nahrwhals_b200.sdgrid import sdgrid, to_solver_inputs
    if seed == 0:
        mat, ly, lx = load_testrun()
    else:
        mat, lx, ly, _ = sdgrid(18 + 3 * seed, [3, 3, 2], chain_depth=2 + seed % 2, seed=seed)
    aln = to_solver_inputs(mat)
    gl = np.concatenate([[0], np.cumsum(lx)])
    depth = 1 + seed % 3
    tsv = oracle.solve(aln, np.asarray(lx, float), np.asarray(ly, float), maxdepth=depth, maxdup=2, maxwidth=200,
                       minreport=0.9, maxreport=300).tsv
    O = ro.format_julia_output(tsv, gl, depth)
    P = nb.format_julia_output(tsv, gl)
    summ = ro.logfile_summary(O["rows"])
    if isinstance(summ["res_ref"], list):
        summ["res_ref"] = ro.NA
    meta = dict(META, y_seqname=None if seed % 2 else "chr1_asm", y_start=None if seed % 2 else "770239",
                y_end=None if seed % 2 else "2005740", compression=None if seed % 2 else 10000)
    log = dict(meta, mut_simulated=P.log["mut_simulated"], mut_tested=P.log["mut_tested"], depth=P.log["depth"],
               exceeds_x=False, exceeds_y=seed == 2, grid_inconsistency=None if seed == 3 else False,
               flip_unsure=P.log["flip_unsure"], cluttered_boundaries=P.log["cluttered_boundaries"])
    want = ro.logfile_row(summ, log)
    got = P.logfile_row(exceeds_x=False, exceeds_y=seed == 2, grid_inconsistency=None if seed == 3 else False, **meta)
    assert got == want
    assert len(got.split("\t")) == len(ro.LOG_COLUMNS) == 37
    assert region.logfile_header() == "\t".join(ro.LOG_COLUMNS)
    f = tmp_path / "nahrwhals_res.tsv"
    region.append_logfile(f, got)
    region.append_logfile(f, got)
    assert f.read_text() == "\t".join(ro.LOG_COLUMNS) + "\n" + got + "\n" + got + "\n"


def test_readme_row_as_logfile_row(oracle):
    """The published row of the test run (README.md:135-139: seqname, start, end, sample, res_ref, res_max, mut_max)
    through the restated region path and the restated save_to_logfile."""
    mat, ly, lx = load_testrun()
    gl = np.concatenate([[0], np.cumsum(lx)])
    A = ro.analyse_region(mat, ly, lx, gl, oracle_solver(oracle), depth=2)
    log = dict(META, exceeds_x=False, exceeds_y=False, grid_inconsistency=False, **A["log"])
    log.setdefault("flip_unsure", False)
    log.setdefault("cluttered_boundaries", False)
    row = dict(zip(ro.LOG_COLUMNS, ro.logfile_row(A, log).split("\t")))
    assert [row[k] for k in ("seqname", "start", "end", "sample", "res_ref", "res_max", "mut_max")] == \
        ["chr1_partial", "1700000", "3300000", "Fasta_x_Fasta_y", "71.378", "99.595", "4_12_inv+11_18_del"]
    assert row["width_orig"] == "1600000" and row["search_depth"] == "2"
    # 4_12_inv: blocks 4..12 of the x axis, offset by the region start
    assert row["mut1_start"] == str(1700000 + int(np.cumsum(lx)[2])) and row["mut2_start"] != "NA" and row["mut3_start"] == "NA"


@pytest.mark.gpu
def test_gpu_readme_row_as_logfile_row(solver, oracle):
    import nahrwhals_b200 as nb
    mat, ly, lx = load_testrun()
    gl = np.concatenate([[0], np.cumsum(lx)])
    P = nb.analyse_region(solver, mat, ly, lx, gl, depth=2)
    A = ro.analyse_region(mat, ly, lx, gl, oracle_solver(oracle), depth=2)
    log = dict(META, exceeds_x=False, exceeds_y=False, grid_inconsistency=False, **A["log"])
    log.setdefault("flip_unsure", False)
    log.setdefault("cluttered_boundaries", False)
    got = P.logfile_row(**META)
    assert got == ro.logfile_row(A, log)
    assert got.startswith("chr1_partial\t1700000\t3300000\tFasta_x_Fasta_y\tNA\tNA\tNA\t1600000\t1\t71.378\t99.595\t")
    assert "\t4_12_inv+11_18_del\t" in got


def _append_worker(args):
    path, tag, n = args
    from nahrwhals_b200 import region
    for k in range(n):
        region.append_logfile(path, "%s\t%d\t%s" % (tag, k, "x" * 3000))     # longer than a pipe buffer's atomic write
    return n


def test_log_append_concurrent_processes(tmp_path):
    """The reference's PSOCK workers append to one nahrwhals_res.tsv under filelock (R/save_logfile.R:197-221): several
    processes appending at once give one header and whole rows only."""
    import multiprocessing as mp
    from nahrwhals_b200 import region
    f = str(tmp_path / "nahrwhals_res.tsv")
    with mp.get_context("fork").Pool(6) as pool:
        done = pool.map(_append_worker, [(f, "w%d" % w, 40) for w in range(6)])
    assert sum(done) == 240
    lines = open(f).read().split("\n")
    assert lines[0] == region.logfile_header() and lines[-1] == "" and len(lines) == 242
    seen = {}
    for ln in lines[1:-1]:
        tag, k, pad = ln.split("\t")
        assert pad == "x" * 3000
        seen.setdefault(tag, []).append(int(k))
    assert sorted(seen) == ["w%d" % w for w in range(6)] and all(v == list(range(40)) for v in seen.values())


def test_r_fixed15_c_abi_matches_restatement():
    """The C side's number formatter (through nw_region_logrow, host only) vs the restatement's on random doubles."""
    import ctypes as C
    from nahrwhals_b200 import region
    L = region._lib_region()
    rng = np.random.default_rng(5)
    vals = list(rng.integers(0, 10 ** 9, 300).astype(float)) + list(np.round(rng.random(300) * 100, 3)) + \
        list(rng.random(300) * 10.0 ** rng.integers(-6, 12, 300)) + [0.0, 1e5, 1e15, 123456789012345.0, 99.9995, 1e-7, 2.5e-5]
    for v in vals:
        S = region.NwRegionSummary()
        S.res_ref = float(v)
        S.res_max = float("nan")
        for k in range(15):
            S.maxres[k] = float("nan")
        m = region.NwLogMeta(b"c", b"0", b"0", b"s", None, None, None, b"1", float("nan"), 0, 0, 0)
        buf = C.create_string_buffer(4096)
        n = C.c_int64()
        assert L.nw_region_logrow(C.byref(S), b"ref", C.byref(m), buf, 4096, C.byref(n)) == 0
        f = buf.value.decode().split("\t")
        assert f[9] == ro.r_fixed15(v) and f[10] == "NA", (v, f[9], ro.r_fixed15(v))

"""Regionfile mode from the alignments on (nahrwhals_b200.pipeline): PAF tables -> rows of nahrwhals_res.tsv, against the
chain of the R restatements (wrapper_paf_to_bitlocus -> wrapper :160-191 -> save_to_logfile)."""
import numpy as np
import pytest

from oracle import nw_bitlocus_oracle as bo
from oracle import nw_region_oracle as ro
from tests.test_bitlocus import sd_paf, _testrun
from tests.test_region_oracle import oracle_solver


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle_lib
    return oracle_lib


def oracle_row(oracle, paf, meta, depth, minlen, compression, chunklen):
    O = bo.wrapper_paf_to_bitlocus(paf, minlen, compression, chunklen)
    base = dict(meta, xpad=meta.get("xpad", "1"), exceeds_x=False, exceeds_y=False)
    if O["status"] == "cluttered":
        summ = dict(res_ref=0.0, res_max=0.0, n_res_max=1, mut_max="ref", maxres={})
        log = dict(base, mut_simulated=0, mut_tested=0, depth=depth, grid_inconsistency=False, flip_unsure=False,
                   cluttered_boundaries=True)
        return ro.logfile_row(summ, log)
    if O["status"] != "ok":
        return None
    g = O["grid"]
    A = ro.analyse_region(g["mat"], g["len_y"], g["len_x"], g["gridlines_x"], oracle_solver(oracle), depth=depth,
                          compression=compression)
    log = dict(base, grid_inconsistency=not g["converged"], **A["log"])
    log.setdefault("flip_unsure", False)
    log.setdefault("cluttered_boundaries", False)
    if isinstance(A["res_ref"], list):
        A["res_ref"] = ro.NA
    return ro.logfile_row(A, log)


def test_empty_row_matches_oracle():
    """The row of a region whose PAF is cluttered (res_empty through save_to_logfile); host only."""
    from nahrwhals_b200 import pipeline
    meta = dict(chr="chr7", start="100000", end="2e+06", samplename="a_b")
    got = pipeline._empty_row(meta, 3, True)
    summ = dict(res_ref=0.0, res_max=0.0, n_res_max=1, mut_max="ref", maxres={})
    log = dict(meta, xpad="1", exceeds_x=False, exceeds_y=False, mut_simulated=0, mut_tested=0, depth=3,
               grid_inconsistency=False, flip_unsure=False, cluttered_boundaries=True)
    assert got == ro.logfile_row(summ, log)
    f = got.split("\t")
    assert f[:4] == ["chr7", "100000", "2e+06", "a_b"] and f[7] == "1900000" and f[9:13] == ["0", "0", "1", "ref"]
    assert f[21] == "TRUE" and f[22:] == ["NA"] * 15


@pytest.mark.gpu
def test_gpu_pipeline_matches_oracle_chain(solver, oracle, tmp_path):
    import nahrwhals_b200 as nb
    pafs = [sd_paf(s, n_blocks=int(12 + 4 * (s % 4))) for s in range(8)]        # seed 3 is a cluttered PAF
    from nahrwhals_b200.sdgrid import matrix_to_chunked_paf
    pafs.insert(5, matrix_to_chunked_paf(np.array([[250000]]), np.array([250000]), np.array([250000])))  # one clean alignment
    pafs.append({k: np.zeros(0) for k in pafs[0]})                                # no alignment: dropped
    metas = [dict(chr="chr%d" % (k + 1), start=str(100000 * k), end=str(100000 * k + 400000), samplename="x_y")
             for k in range(len(pafs))]
    f = tmp_path / "nahrwhals_res.tsv"
    rows, det = nb.run_regions(solver, pafs, metas, depth=2, minlen=1000.0, compression=1000.0, chunklen=10000.0, logfile=f)
    want = [oracle_row(oracle, p, m, 2, 1000.0, 1000.0, 10000.0) for p, m in zip(pafs, metas)]
    assert rows == want
    assert rows[-1] is None and sum(r is not None for r in rows) == 9
    assert rows[5].split("\t")[9:13] == ["100", "100", "1", "ref"]     # a 1 x 1 gridmatrix: 100 % in reference state
    assert any(b.status == "cluttered" for b, _ in det) and any(r is not None and r.solver_called for _, r in det)
    lines = f.read_text().split("\n")
    assert lines[0] == "\t".join(ro.LOG_COLUMNS) and lines[1:-1] == [r for r in rows if r is not None]
    # the same as one call of the C ABI (nw_regions_run)
    f2 = tmp_path / "res2.tsv"
    rows2, st = nb.run_regions_native(solver, pafs, metas, 1000.0, 1000.0, 10000.0, depth=2, logfile=f2)
    assert rows2 == want and f2.read_text() == f.read_text()
    for (b, _), r, s_ in zip(det, rows, st):
        assert (s_ == 1) == (b.status == "cluttered") and (s_ < 0) == (r is None)
    assert st[-1] < 0 and st[3] == 1
    # two contexts sharing batches of three regions
    s2 = nb.Solver(0)
    try:
        f3 = tmp_path / "res3.tsv"
        rows3, st3 = nb.run_regions_native([solver, s2], pafs, metas, 1000.0, 1000.0, 10000.0, depth=2, logfile=f3,
                                           regions_per_batch=3)
        assert rows3 == want and st3 == st and f3.read_text() == f.read_text()
        # and from PAF files (one per region; a missing file drops its region only)
        from tests.test_paf import write_chunk_paf
        paths = []
        for k, p in enumerate(pafs):
            paths.append(tmp_path / ("r%d.paf" % k))
            write_chunk_paf(paths[-1], p, tags=bool(k % 2))
        paths[2] = tmp_path / "missing.paf"
        rows4, st4 = nb.run_region_files([solver, s2], paths, metas, 1000.0, 1000.0, 10000.0, depth=2, regions_per_batch=4)
        assert st4[2] == -2 and rows4[2] is None
        assert [r for k, r in enumerate(rows4) if k != 2] == [r for k, r in enumerate(want) if k != 2]
    finally:
        s2.close()


@pytest.mark.gpu
def test_gpu_pipeline_test_run_row(solver, oracle):
    """The bundled test run from its chunked alignments on, with the reference's default parameters, at the depth the
    published row was made with: chr1_partial 1700000 3300000 ... 71.378 99.595 ... 4_12_inv+11_18_del."""
    import nahrwhals_b200 as nb
    mat, lx, ly, ap, chunks = _testrun()
    meta = dict(chr="chr1_partial", start="1700000", end="3300000", samplename="Fasta_x_Fasta_y")
    rows, _ = nb.run_regions(solver, [chunks], [meta], depth=2)
    assert rows[0] == oracle_row(oracle, chunks, meta, 2, ap["minlen"], ap["compression"], ap["chunklen"])
    row = dict(zip(ro.LOG_COLUMNS, rows[0].split("\t")))
    assert [row[k] for k in ("seqname", "start", "end", "sample", "res_ref", "res_max", "mut_max")] == \
        ["chr1_partial", "1700000", "3300000", "Fasta_x_Fasta_y", "71.378", "99.595", "4_12_inv+11_18_del"]
    rows2, st = nb.run_regions_native(solver, [chunks], [meta], ap["minlen"], ap["compression"], ap["chunklen"], depth=2)
    assert rows2 == rows and st == [0]

"""GPU parity at the sizes bench.py runs (BASELINE.json configs 2, 3, 5 and the north_star S150 depth-3 target):
the CUDA path through the C ABI against fixtures the CPU oracle produced ONCE for exactly these workloads
(tests/golden/fullsize_*.json, written by tests/golden/make_fullsize_fixtures.py: minutes of CPU per search, so the
oracle itself cannot run inside the test).  Compared: per-level generated / scored / DP-cell counters
(solver.jl:589-642, :536), number of ids, best score, SHA-256 over every id's discovery level, DP terminal cost, total
and Float32 score bits, over every provenance edge, and over the report TSV (solver.jl:747-753)."""
import hashlib
import importlib.util
import json
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _fx():
    spec = importlib.util.spec_from_file_location("make_fullsize_fixtures", os.path.join(GOLD, "make_fullsize_fixtures.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def check_against_fixture(solver, name, **extra):
    fxm = _fx()
    F = json.load(open(os.path.join(GOLD, "fullsize_%s.json" % name)))
    sys.path.insert(0, ROOT)
    import bench
    aln, lx, ly, flags, desc = bench.make_workload(name, 1)
    assert desc == F["workload"] and (aln.shape[0], aln.shape[1]) == (F["n_x"], F["n_y"])
    G = solver.solve(aln, lx, ly, keep_ids=True, **flags, **extra)
    assert G.generated == F["generated"], (G.generated, F["generated"])
    assert G.scored == F["scored"], (G.scored, F["scored"])
    assert G.cells == F["cells"], (G.cells, F["cells"])
    assert int(G.stats.n_ids) == F["n_ids"] and G.stats.best == F["best"]
    ids = G.ids
    got = fxm.id_digests(ids["level"], ids["cost_bp"], ids["total_bp"], ids["score"])
    assert got == F["ids"], "per-id arrays differ from the oracle's"
    ed = G.edges
    assert len(ed["child"]) == F["n_edges"]
    assert fxm.edge_digest(ed["child"], ed["parent"], ed["moves"]) == F["edges"], "provenance edges differ from the oracle's"
    assert len(G.scores) == F["n_trajs"]
    assert G.tsv.split("\n")[:3] == F["tsv_head"]
    assert hashlib.sha256(G.tsv.encode()).hexdigest() == F["tsv_sha256"], "report TSV differs from the oracle's"
    return G


def test_s150_beam_depth4_matches_oracle_fixture(solver):
    """BASELINE config 3 on one GPU: 150 blocks, -d 4 -u 2 -w 1000."""
    check_against_fixture(solver, "s150")


def test_s150_exhaustive_depth3_matches_oracle_fixture(solver):
    """north_star target workload: 150 blocks, exhaustive depth 3 (6.1 M candidates, 3.8 M scored)."""
    check_against_fixture(solver, "s150d3")


def test_s64_exhaustive_depth3_matches_oracle_fixture(solver):
    """BASELINE config 2 = the bench's headline workload: 17.6 M candidates, 10.8 M scored, every id and edge."""
    check_against_fixture(solver, "s64")


def test_s64_exhaustive_unverified_dedup(solver):
    """The same search with the element-wise dedup verifier switched off (NW_FLAG_NO_VERIFY): identical results."""
    check_against_fixture(solver, "s64", verify_dedup=False)


def test_genome_sample_matches_oracle_fixture(solver):
    """BASELINE config 5: 200 windows of the whole-genome plan bench.py --workload genome runs, through the solver stage
    (nw_region_analyse_many) against the R restatement around the oracle solver."""
    import nahrwhals_b200 as nb
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from region_batch import genome_plan, make_genome_regions
    fxm = _fx()
    F = json.load(open(os.path.join(GOLD, "fullsize_genome.json")))
    plan = genome_plan(F["plan"]["n"], seed=F["plan"]["seed"])
    assert len(plan) == F["plan"]["windows_after_blacklist"]
    idx = [w["window"] for w in F["windows"]]
    regs = make_genome_regions(plan, idx)
    res = nb.analyse_regions(solver, regs, depth=3, maxdup=2, init_width=1000, minreport=0.98, regions_per_batch=64)
    assert len(res) == len(idx)
    for w, P, (mat, ly, lx, gl) in zip(F["windows"], res, regs):
        assert (mat.shape[1], mat.shape[0]) == (w["n_x"], w["n_y"])
        assert P.mut_max == w["mut_max"], (w["window"], P.mut_max, w["mut_max"])
        assert (P.res_ref, P.res_max, P.n_res_max) == (w["res_ref"], w["res_max"], w["n_res_max"]), w["window"]
        assert bool(P.solver_called) == w["solver_called"] and len(P.rows) == w["n_rows"]
        assert hashlib.sha256(fxm.canon_rows(P.rows).encode()).hexdigest() == w["rows_sha256"], w["window"]

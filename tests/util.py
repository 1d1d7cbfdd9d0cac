"""Shared parity helpers: compare a GPU search (nahrwhals_b200.Solver) with the oracle id by id."""
import numpy as np


def assert_search_equal(G, O, check_tsv=True):
    """G: SolveResult with keep_ids; O: oracle_lib.OracleResult (full)."""
    nl = G.stats.n_levels
    gen = [int(v) for v in G.stats.generated[:nl]]
    sco = [int(v) for v in G.stats.scored[:nl]]
    cel = [int(v) for v in G.stats.cells[:nl]]
    assert gen == [int(v) for v in O.generated], (gen, O.generated)
    assert sco == [int(v) for v in O.scored], (sco, O.scored)
    assert cel == [int(v) for v in O.cells], (cel, O.cells)
    assert int(G.stats.n_ids) == O.n_ids
    assert G.stats.best == O.best
    ids = G.ids
    # discovery level of every id
    assert np.array_equal(ids["level"], O.level)
    # Float32 scores bit-exact (incl. -inf)
    assert np.array_equal(ids["score"].view(np.uint32), O.score.view(np.uint32))
    # integer DP terminal cost: oracle NaN = early exit (-1), inf = unreachable (-2)
    ocost = np.where(np.isnan(O.cost), -1.0, np.where(np.isinf(O.cost), -2.0, O.cost)).astype(np.int64)
    assert np.array_equal(ids["cost_bp"], ocost)
    assert np.array_equal(ids["total_bp"], O.total.astype(np.int64))
    # provenance edges: the oracle stores them per child in insertion order; the GPU emits them in discovery order
    ed = G.edges
    order = np.argsort(ed["child"], kind="stable")
    assert np.array_equal(ed["child"][order], np.repeat(np.arange(1, O.n_ids + 1), np.diff(O.edge_off)))
    assert np.array_equal(ed["parent"][order], O.edge_parent)
    assert np.array_equal(ed["moves"][order], O.edge_move)
    if check_tsv:
        assert G.tsv == O.tsv

"""PAF preparation: the R restatement on the CPU, and the GPU path (include/nwpaf.h) against it."""
import numpy as np
import pytest

from oracle import nw_paf_oracle as po


def chunked_paf(rng, n_segments=6, chunk=10000, jitter=0, span=400000, drop=0.0, extra=0):
    """Alignments as chunked minimap2 reports them: every true segment cut into chunk-long pieces (optionally with a few
    bp of jitter at the cuts, dropped pieces and unrelated short hits)."""
    rows = []
    for _ in range(n_segments):
        length = min(chunk * int(rng.integers(1, 9)) + int(rng.integers(0, chunk)), span // 2)
        t0 = int(rng.integers(0, span - length))
        q0 = int(rng.integers(0, span - length))
        strand = 1 if rng.random() < 0.6 else -1
        for k in range(0, length, chunk):
            ln = min(chunk, length - k)
            if rng.random() < drop:
                continue
            j0 = int(rng.integers(-jitter, jitter + 1)) if jitter else 0
            j1 = int(rng.integers(-jitter, jitter + 1)) if jitter else 0
            ts, te = t0 + k + j0, t0 + k + ln + j1
            if strand > 0:
                qs, qe = q0 + k + j0, q0 + k + ln + j1
            else:
                qs, qe = q0 + length - k - ln - j1, q0 + length - k - j0
            if te - ts < 50 or qe - qs < 50:
                continue
            rows.append((chunk, qs, qe, strand, span, ts, te, te - ts - int(rng.integers(0, 30)), te - ts, 60))
    for _ in range(extra):
        ln = int(rng.integers(200, 3000))
        t0, q0 = int(rng.integers(0, span - ln)), int(rng.integers(0, span - ln))
        rows.append((chunk, q0, q0 + ln, 1 if rng.random() < 0.5 else -1, span, t0, t0 + ln, ln - 5, ln, 30))
    order = rng.permutation(len(rows))
    a = np.array([rows[k] for k in order], dtype=np.float64).reshape(-1, 10)
    return {k: a[:, i].copy() for i, k in enumerate(po.COLS)}


def test_oracle_stitches_chunks():
    """Chunks of one alignment are stitched back (forward and reverse strand), unrelated alignments are left alone."""
    rng = np.random.default_rng(0)
    p = chunked_paf(rng, n_segments=1, chunk=10000)
    out = po.compress_paf_fnct(po.swap_minus(p), second_run=True, chunklen=10000, compression=1000)
    assert len(out["tstart"]) == 1
    assert out["tend"][0] - out["tstart"][0] == p["alen"].sum() == out["alen"][0]
    # the full preparation ends with slope-one alignments on the compression grid
    p = chunked_paf(np.random.default_rng(3), n_segments=5, jitter=3, extra=4)
    out = po.load_and_prep_paf(p, minlen=2000, compression=1000)
    assert len(out["tstart"]) >= 3
    assert np.all(np.abs(out["tend"] - out["tstart"]) == np.abs(out["qend"] - out["qstart"]))
    assert np.all(out["tstart"] % 1000 == 0) and np.all(out["qstart"] % 1000 == 0)
    neg = out["strand"] < 0
    assert np.all(out["qstart"][neg] > out["qend"][neg]) and np.all(out["qstart"][~neg] < out["qend"][~neg])


def _same(P, O):
    for k in po.COLS:
        assert np.array_equal(P[k], O[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(12))
def test_gpu_paf_matches_oracle(solver, seed):
    import nahrwhals_b200 as nb
    rng = np.random.default_rng(100 + seed)
    p = chunked_paf(rng, n_segments=int(rng.integers(1, 14)), chunk=int(rng.choice([5000, 10000])),
                    jitter=int(rng.choice([0, 0, 2, 30, 400])), span=int(rng.choice([40000, 300000, 900000])),
                    drop=float(rng.choice([0.0, 0.1])), extra=int(rng.integers(0, 12)))
    chunklen = float(p["qlen"][0]) if len(p["qlen"]) else 10000.0
    for second_run, compression in ((True, 1000.0), (False, 1000.0), (True, 10000.0)):
        O = po.compress_paf_fnct(po.swap_minus(p), second_run=second_run, chunklen=chunklen, compression=compression)
        P = nb.compress_paf(solver, po.swap_minus(p), second_run=second_run, chunklen=chunklen, compression=compression)
        _same(P, O)
    for minlen, compression in ((500.0, 1000.0), (3000.0, 5000.0)):
        O = po.load_and_prep_paf(p, minlen, compression, chunklen)
        P = nb.load_and_prep_paf(solver, p, minlen, compression, chunklen)
        _same(P, O)
    # quadrant count given explicitly (R/make_whole_genome_list.R:105 passes 10)
    O = po.compress_paf_fnct(po.swap_minus(p), second_run=False, chunklen=chunklen, n_quadrants_per_axis=10)
    P = nb.compress_paf(solver, po.swap_minus(p), second_run=False, chunklen=chunklen, n_quadrants_per_axis=10)
    _same(P, O)


@pytest.mark.gpu
def test_gpu_paf_edge_cases_and_chain(solver):
    """Empty / single-row tables, bad input, and the whole native chain PAF -> grid -> region stage."""
    import nahrwhals_b200 as nb
    empty = {k: np.zeros(0) for k in po.COLS}
    assert len(nb.load_and_prep_paf(solver, empty, 100, 1000)["tstart"]) == 0
    one = chunked_paf(np.random.default_rng(1), n_segments=1, chunk=100000)
    one = {k: v[:1] for k, v in one.items()}
    _same(nb.compress_paf(solver, one, second_run=True), po.compress_paf_fnct(one, second_run=True, compression=1000.0))
    bad = dict(one)
    bad["strand"] = np.array([0.0])
    with pytest.raises(Exception):
        nb.compress_paf(solver, bad)
    p = chunked_paf(np.random.default_rng(9), n_segments=7, jitter=2, span=200000)
    T = nb.load_and_prep_paf(solver, p, 1000, 1000)
    assert T["_gpu_launches"] >= 1 and T["_pairs_found"] > 0
    G = nb.paf_to_gridmatrix(solver, T["tstart"], T["tend"], T["qstart"], T["qend"])
    assert G.mat.shape[0] >= 2 and G.mat.shape[1] >= 2
    R = nb.analyse_region(solver, G.mat, G.len_y, G.len_x, G.gridlines_x, depth=2)
    assert R.res_ref is not None and R.res_max >= R.res_ref


@pytest.mark.gpu
def test_gpu_paf_batch_matches_oracle(solver):
    """nw_paf_prepare_many: tables of very different sizes (incl. empty, single-row and a malformed one) in one launch."""
    import nahrwhals_b200 as nb
    rng = np.random.default_rng(77)
    tabs = [chunked_paf(rng, n_segments=int(rng.integers(1, 40)), chunk=int(rng.choice([5000, 10000])),
                        jitter=int(rng.choice([0, 2, 30])), span=int(rng.choice([40000, 300000, 2000000])),
                        drop=float(rng.choice([0.0, 0.1])), extra=int(rng.integers(0, 12))) for _ in range(40)]
    tabs.insert(3, {k: np.zeros(0) for k in po.COLS})
    tabs.insert(9, {k: v[:1] for k, v in tabs[0].items()})
    bad = {k: v.copy() for k, v in tabs[1].items()}
    bad["strand"][0] = 0.0
    tabs.insert(5, bad)
    for minlen, compression in ((500.0, 1000.0), (3000.0, 5000.0)):
        got, rcs = nb.load_and_prep_pafs(solver, tabs, minlen, compression, 10000.0)
        for k, (t, rc) in enumerate(zip(got, rcs)):
            if k == 5:
                assert t is None and rc == -1
                continue
            assert rc == 0
            _same(t, po.load_and_prep_paf(tabs[k], minlen, compression, 10000.0))
        assert max(t["_gpu_launches"] for t in got if t is not None) == 1


@pytest.mark.gpu
def test_gpu_paf_pair_segment_overflow(solver):
    """Far more pairs than alignments (40 near-identical short alignments: every ordered pair qualifies): the table's
    output segment overflows and the launch is repeated with the exact size."""
    import nahrwhals_b200 as nb
    n = 40
    rng = np.random.default_rng(4)
    j = rng.integers(0, 5, size=(n, 4)).astype(np.float64)
    p = dict(qlen=np.full(n, 1000.0), qstart=1000.0 + j[:, 0], qend=1010.0 + j[:, 1], strand=np.ones(n),
             tlen=np.full(n, 1e6), tstart=2000.0 + j[:, 2], tend=2010.0 + j[:, 3], nmatch=np.full(n, 150.0),
             alen=np.full(n, 10.0), mapq=np.full(n, 60.0))
    O = po.compress_paf_fnct(p, second_run=False, chunklen=1000.0, compression=1000.0)
    P = nb.compress_paf(solver, p, second_run=False, chunklen=1000.0, compression=1000.0)
    assert P["_pairs_found"] > 2 * n + 64 and P["_gpu_launches"] == 2
    _same(P, O)
    other = chunked_paf(np.random.default_rng(2), n_segments=5)
    got, rcs = nb.load_and_prep_pafs(solver, [other, p, other], 5.0, 10.0, 1000.0)
    assert rcs == [0, 0, 0]
    for t, src in zip(got, (other, p, other)):
        _same(t, po.load_and_prep_paf(src, 5.0, 10.0, 1000.0))


# ---------------------------------------------------------------------------------------------------------------
# file mode: compress_paf_fnct(inpaf_link, outpaf_link) -- R/seqbuilder_functions.R:107
# ---------------------------------------------------------------------------------------------------------------
def write_chunk_paf(path, p, tags=False, dup_lines=0):
    """A PAF file as correct_paf leaves it: qname = <seq>_<start>-<end> of the chunk, 12 columns (optionally tags)."""
    lines = []
    for k in range(len(p["qstart"])):
        qn = "chr1_partial_%d-%d" % (int(p["qstart"][k]), int(p["qend"][k]) - 1)
        f = [qn, int(p["qlen"][k]), int(p["qstart"][k]), int(p["qend"][k]), "+" if p["strand"][k] > 0 else "-", "chr1_t",
             int(p["tlen"][k]), int(p["tstart"][k]), int(p["tend"][k]), int(p["nmatch"][k]), int(p["alen"][k]), int(p["mapq"][k])]
        if tags:
            f += ["tp:A:P", "cm:i:%d" % (k % 7)]
        lines.append("\t".join(str(v) for v in f))
    lines += lines[:dup_lines]                                # exact duplicates: unique() drops them
    open(path, "w").write("\n".join(lines) + "\n")


def test_r_number_formatting():
    """as.character / write.table of doubles as R prints them (15 significant digits, scientific when narrower)."""
    for v, want in [(100000, "1e+05"), (123456, "123456"), (1e6, "1e+06"), (1200000, "1200000"), (12000000, "1.2e+07"),
                    (0.5, "0.5"), (1e-4, "1e-04"), (0.0001234, "0.0001234"), (99999, "99999"), (0.1 + 0.2, "0.3"),
                    (1 / 3, "0.333333333333333"), (1e15, "1e+15"), (123456789012, "123456789012"), (-100000, "-1e+05"),
                    (3e5, "3e+05"), (150.0, "150"), (0, "0")]:
        assert po.r_format_double(v) == want, v
    assert po.r_format(100000.0, True) == "100000"


def test_read_paf_file_c_abi(tmp_path):
    """nw_paf_read_file is host-only: a PAF with tags, against the restatement's reader; malformed files fail."""
    import nahrwhals_b200 as nb
    p = chunked_paf(np.random.default_rng(3), n_segments=4, jitter=2)
    f = tmp_path / "in.paf"
    write_chunk_paf(f, p, tags=True)
    got = nb.read_paf(f)
    want = po.read_paf_file(str(f))[0]
    for k in po.COLS:
        assert np.array_equal(got[k], want[k]) and np.array_equal(got[k], p[k]), k
    (tmp_path / "bad.paf").write_text("a\t1\t2\n")
    with pytest.raises(Exception):
        nb.read_paf(tmp_path / "bad.paf")
    with pytest.raises(Exception):
        nb.read_paf(tmp_path / "missing.paf")


def test_oracle_file_mode_names(tmp_path):
    """Chunks of a forward alignment melt into one row whose qname is rebuilt from the first chunk's start and the last
    chunk's end (R/compress_paf.R:49-59), with R's double formatting of `qend2 - 1`."""
    n = 12
    p = dict(qlen=np.full(n, 10000.0), qstart=10000.0 * np.arange(n), qend=10000.0 * (np.arange(n) + 1), strand=np.ones(n),
             tlen=np.full(n, 5e6), tstart=500000 + 10000.0 * np.arange(n), tend=500000 + 10000.0 * (np.arange(n) + 1),
             nmatch=np.full(n, 9990.0), alen=np.full(n, 10000.0), mapq=np.full(n, 60.0))
    write_chunk_paf(tmp_path / "in.paf", p, dup_lines=3)
    nin, nout = po.compress_paf_file(str(tmp_path / "in.paf"), str(tmp_path / "out.paf"), 10000.0)
    assert (nin, nout) == (n, 1)
    f = (tmp_path / "out.paf").read_text().strip().split("\t")
    assert f[0] == "chr1_partial_0-119999" and f[2:4] == ["0", "120000"] and f[7:9] == ["500000", "620000"]
    assert f[9:11] == ["119880", "120000"] and len(f) == 12     # nmatch / alen are doubles after the merge: 120000 stays fixed


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(8))
def test_gpu_paf_file_mode_matches_oracle(solver, tmp_path, seed):
    import nahrwhals_b200 as nb
    rng = np.random.default_rng(500 + seed)
    if seed == 0:     # one long forward alignment whose merged nmatch / alen / name hit R's scientific notation
        n = 10
        p = dict(qlen=np.full(n, 10000.0), qstart=1.0 + 10000.0 * np.arange(n), qend=1.0 + 10000.0 * (np.arange(n) + 1),
                 strand=np.ones(n), tlen=np.full(n, 5e6), tstart=500000 + 10000.0 * np.arange(n),
                 tend=500000 + 10000.0 * (np.arange(n) + 1), nmatch=np.full(n, 10000.0), alen=np.full(n, 10000.0),
                 mapq=np.full(n, 60.0))
    else:
        p = chunked_paf(rng, n_segments=int(rng.integers(1, 12)), chunk=10000, jitter=int(rng.choice([0, 2, 30])),
                        span=int(rng.choice([40000, 300000, 900000])), drop=float(rng.choice([0.0, 0.1])),
                        extra=int(rng.integers(0, 6)))
        if seed == 1:
            p = {k: v[:1] for k, v in p.items()}
    fin, fo, fg = tmp_path / "in.paf", tmp_path / "oracle.paf", tmp_path / "gpu.paf"
    write_chunk_paf(fin, p, tags=(seed % 3 == 2), dup_lines=seed % 4)
    want = po.compress_paf_file(str(fin), str(fo), 10000.0)
    got = nb.compress_paf_file(solver, fin, fg, 10000.0)
    assert got == want
    assert fg.read_text() == fo.read_text()
    if seed == 0:
        assert "1e+05" in fg.read_text()
    # the molten file is what load_and_prep_paf_for_gridplot_transform reads next
    T = nb.read_paf(fg)
    O = po.read_paf_file(str(fo))[0]
    for k in po.COLS:
        assert np.array_equal(T[k], O[k]), k


def write_multi_qname_paf(path, tables, names):
    """A chunked all-vs-all PAF: several query sequences, chunk qnames <seq>_<start>-<end>, rows interleaved."""
    lines = []
    for p, nm in zip(tables, names):
        for k in range(len(p["qstart"])):
            f = ["%s_%d-%d" % (nm, int(p["qstart"][k]), int(p["qend"][k]) - 1), int(p["qlen"][k]), int(p["qstart"][k]),
                 int(p["qend"][k]), "+" if p["strand"][k] > 0 else "-", "chr%d" % (1 + k % 2), int(p["tlen"][k]),
                 int(p["tstart"][k]), int(p["tend"][k]), int(p["nmatch"][k]), int(p["alen"][k]), int(p["mapq"][k])]
            lines.append("\t".join(str(v) for v in f))
    order = np.random.default_rng(len(lines)).permutation(len(lines))
    open(path, "w").write("\n".join(lines[k] for k in order) + "\n")


def test_oracle_qname_mode(tmp_path):
    """Per query sequence: rows of other sequences never merge, names lose the chunk suffix and stay reduced, output appended."""
    n = 6
    mk = lambda off: dict(qlen=np.full(n, 10000.0), qstart=10000.0 * np.arange(n), qend=10000.0 * (np.arange(n) + 1),  # noqa: E731
                          strand=np.ones(n), tlen=np.full(n, 5e6), tstart=off + 10000.0 * np.arange(n),
                          tend=off + 10000.0 * (np.arange(n) + 1), nmatch=np.full(n, 9990.0), alen=np.full(n, 10000.0),
                          mapq=np.full(n, 60.0))
    write_multi_qname_paf(tmp_path / "in.paf", [mk(100000), mk(100000)], ["ctgA", "ctg_B"])   # identical coordinates
    (tmp_path / "out.paf").write_text("previous line\n")
    nin, nout = po.compress_paf_file_by_qname(str(tmp_path / "in.paf"), str(tmp_path / "out.paf"), 10000.0)
    lines = (tmp_path / "out.paf").read_text().strip().split("\n")
    assert (nin, nout) == (12, 2) and lines[0] == "previous line" and len(lines) == 3
    assert sorted(ln.split("\t")[0] for ln in lines[1:]) == ["ctgA", "ctg_B"]
    assert all(ln.split("\t")[2:4] == ["0", "60000"] for ln in lines[1:])


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(4))
def test_gpu_paf_qname_mode_matches_oracle(solver, tmp_path, seed):
    import nahrwhals_b200 as nb
    rng = np.random.default_rng(900 + seed)
    tabs = [chunked_paf(rng, n_segments=int(rng.integers(1, 10)), chunk=10000, jitter=int(rng.choice([0, 2, 30])),
                        span=int(rng.choice([300000, 900000])), drop=float(rng.choice([0.0, 0.1])),
                        extra=int(rng.integers(0, 5))) for _ in range(int(rng.integers(1, 7)))]
    if seed == 1:
        tabs.append({k: v[:1] for k, v in tabs[0].items()})            # a query sequence with one alignment
    names = ["h1tg%06dl" % k if k % 2 else "contig_%d" % k for k in range(len(tabs))]
    fin, fo, fg = tmp_path / "in.paf", tmp_path / "oracle.paf", tmp_path / "gpu.paf"
    write_multi_qname_paf(fin, tabs, names)
    want = po.compress_paf_file_by_qname(str(fin), str(fo), 10000.0, 10)
    got = nb.compress_paf_file_by_qname(solver, fin, fg, 10000.0, 10)
    assert got == want and fg.read_text() == fo.read_text()
    assert nb.compress_paf_file_by_qname(solver, fin, fg, 10000.0, 10) == want       # appended, not truncated
    assert fg.read_text() == fo.read_text() * 2

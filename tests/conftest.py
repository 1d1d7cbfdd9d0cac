import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    # a fresh checkout has no built artefacts (they are git-ignored): build the product library, the oracle and the CPU
    # harnesses once (nvcc cross-compiles sm_100a without a GPU), exactly as __graft_entry__.build() does
    lib = os.path.join(ROOT, "nahrwhals_b200", "csrc", "libnwsolve.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def solver():
    import nahrwhals_b200 as nb
    s = nb.Solver(0)
    yield s
    s.close()


@pytest.fixture(scope="session")
def testrun():
    """The reference's bundled test-run gridmatrix (tests/golden/make_testrun_fixture.py)."""
    import numpy as np
    g = os.path.join(ROOT, "tests", "golden")
    mat = np.loadtxt(os.path.join(g, "testrun_grid_mut.tsv"), dtype=np.float64, ndmin=2)
    lens = open(os.path.join(g, "testrun_grid_mut_lens.tsv")).read().strip().split("\n")
    ly = np.array(lens[0].split("\t"), dtype=np.int64)
    lx = np.array(lens[1].split("\t"), dtype=np.int64)
    aln = np.ascontiguousarray(np.sign(mat).T.astype(np.int8))
    return aln, lx, ly

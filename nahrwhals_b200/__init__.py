"""nahrwhals_b200 -- B200-native (sm_100a) implementation of NAHRwhals' mutation-space solver.

Only the hot path of the reference (its Julia script solver.jl, called from R/juliasolver.R) lives here:
`csrc/` holds the CUDA kernels and the C ABI (include/nwsolve.h); `solver.py` mirrors the reference's solver
interface on top of that ABI.  There is no CPU fallback.
"""
from .solver import (Solver, SolveResult, read_problem, solve_batch, solve_many, solve_files_many, main,  # noqa: F401
                     MOVE_NAMES)  # noqa: F401
from .region import solve_mutation_depth1, format_julia_output, analyse_region, analyse_regions, RegionResult  # noqa: F401
from .grid import build_grids, paf_to_gridmatrix, GridResult  # noqa: F401
from . import bitlocus, grid, paf, region  # noqa: F401
from .bitlocus import paf_to_bitlocus, auto_params  # noqa: F401
from .pipeline import run_regions, run_regions_native, run_region_files  # noqa: F401
from .paf import compress_paf, load_and_prep_paf, load_and_prep_pafs, compress_paf_file, compress_paf_file_by_qname, read_paf  # noqa: F401
from .sdgrid import sdgrid, to_solver_inputs, write_tsvs, matrix_to_alignments  # noqa: F401

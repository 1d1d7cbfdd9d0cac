"""ctypes binding of libnwsolve.so (include/nwsolve.h).  No CPU fallback: import fails loudly if the CUDA
library has not been built (run `python -c "import __graft_entry__ as g; g.build()"` or `make -C nahrwhals_b200/csrc`)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NWSOLVE_LIB") or os.path.join(HERE, "csrc", "libnwsolve.so")  # NWSOLVE_LIB: another build of the same library


class NwOpts(C.Structure):
    _fields_ = [("maxdepth", C.c_int32), ("maxdup", C.c_int32), ("maxwidth", C.c_int64), ("minreport", C.c_double),
                ("maxreport", C.c_int64), ("keep_ids", C.c_int32), ("flags", C.c_int32)]


class NwStats(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("kernel_launches", C.c_int32), ("generated", C.c_int64 * 16),
                ("scored", C.c_int64 * 16), ("cells", C.c_int64 * 16), ("frontier", C.c_int64 * 16),
                ("level_ms", C.c_float * 16), ("total_ms", C.c_float), ("n_ids", C.c_int64), ("best", C.c_double),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("score_ms", C.c_float),
                ("score_launches", C.c_int32), ("host_ms", C.c_float * 4)]


DIST_ID_BYTES = 128
FLAG_GENERIC_DP = 1
FLAG_NO_TRAJ = 2
FLAG_PAIR_EXPAND = 4
FLAG_VERIFY_DEDUP = 8
FLAG_DEBUG_WEAK_FP = 16
FLAG_NO_VERIFY = 32
FLAG_DEBUG_WEAK_FP_ONCE = 64
FLAG_DP32 = 128

# every symbol include/nwsolve.h declares (tests check that the library exports exactly these)
SYMBOLS = ["nw_default_opts", "nw_last_error", "nw_version", "nw_ctx_create", "nw_ctx_destroy", "nw_solve",
           "nw_solve_files", "nw_result_count", "nw_result_get", "nw_result_cost", "nw_result_arrays", "nw_result_move_count",
           "nw_result_tsv", "nw_result_stats",
           "nw_result_ids", "nw_result_edge_count", "nw_result_edges", "nw_result_free", "nw_score_batch",
           "nw_moveset", "nw_bench_peaks", "nw_read_problem", "nw_free", "nw_dist_unique_id", "nw_dist_init",
           "nw_dist_info", "nw_dist_shutdown", "nw_solve_sharded", "nw_solve_sharded_ctxs", "nw_solve_files_sharded", "nw_device_count", "nw_solve_batch", "nw_solve_many",
           "nw_solve_many_ctxs", "nw_solve_files_many", "nw_results_sizes", "nw_results_gather"]

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("nahrwhals_b200: %s is missing -- the CUDA extension is not built and there is no CPU "
                          "fallback.  Build it with `make -C nahrwhals_b200/csrc` (needs nvcc, sm_100a)." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.nw_default_opts.argtypes = [C.POINTER(NwOpts)]
    L.nw_last_error.restype = C.c_char_p
    L.nw_version.restype = C.c_char_p
    L.nw_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.nw_ctx_destroy.argtypes = [vp]
    L.nw_solve.argtypes = [vp, vp, i32, i32, vp, vp, C.POINTER(NwOpts), C.POINTER(vp)]
    L.nw_solve_files.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(NwOpts), C.POINTER(NwStats)]
    L.nw_result_count.restype = i64
    L.nw_result_count.argtypes = [vp]
    L.nw_result_get.argtypes = [vp, i64, C.POINTER(C.c_float), C.POINTER(i32), C.POINTER(C.POINTER(i32))]
    L.nw_result_cost.argtypes = [vp, i64, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    L.nw_result_arrays.argtypes = [vp] * 8
    L.nw_result_move_count.restype = i64
    L.nw_result_move_count.argtypes = [vp]
    L.nw_result_tsv.argtypes = [vp, C.POINTER(vp), C.POINTER(i64)]
    L.nw_result_stats.argtypes = [vp, C.POINTER(NwStats)]
    L.nw_result_ids.argtypes = [vp, vp, vp, vp, vp]
    L.nw_result_edge_count.restype = i64
    L.nw_result_edge_count.argtypes = [vp]
    L.nw_result_edges.argtypes = [vp, vp, vp, vp]
    L.nw_result_free.argtypes = [vp]
    L.nw_score_batch.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, i64, i32, i32, vp, vp, vp]
    L.nw_moveset.argtypes = [vp, vp, i32, i32, vp, vp]
    L.nw_bench_peaks.argtypes = [vp, C.POINTER(C.c_double * 4)]
    L.nw_read_problem.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(vp), C.POINTER(i32), C.POINTER(i32),
                                  C.POINTER(vp), C.POINTER(vp)]
    L.nw_free.argtypes = [vp]
    L.nw_dist_unique_id.argtypes = [vp]
    L.nw_dist_init.argtypes = [vp, vp, i32, i32]
    L.nw_dist_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
    L.nw_dist_shutdown.argtypes = [vp]
    L.nw_solve_sharded.argtypes = [vp, vp, i32, i32, vp, vp, C.POINTER(NwOpts), C.POINTER(vp)]
    L.nw_solve_sharded_ctxs.argtypes = [vp, i32, vp, i32, i32, vp, vp, C.POINTER(NwOpts), vp]
    L.nw_solve_batch.argtypes = [vp, i32, vp, i64, C.POINTER(NwOpts), vp, vp]
    L.nw_solve_many.argtypes = [vp, vp, i64, C.POINTER(NwOpts), i32, vp]
    L.nw_results_sizes.argtypes = [vp, i64, vp, vp, vp]
    L.nw_results_gather.argtypes = [vp, i64] + [vp] * 9
    L.nw_solve_files_many.argtypes = [vp, i32, vp, vp, vp, i64, C.POINTER(NwOpts), i32, vp]
    L.nw_solve_many_ctxs.argtypes = [vp, i32, vp, i64, C.POINTER(NwOpts), i32, vp]
    _lib = L
    return L


class NwProblem(C.Structure):
    _fields_ = [("aln_sign", C.c_void_p), ("n_x", C.c_int32), ("n_y", C.c_int32), ("len_x", C.c_void_p),
                ("len_y", C.c_void_p)]


class NwError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("nwsolve error %d: %s" % (code, msg))
        self.code = code


def check(rc):
    if rc != 0:
        raise NwError(rc, load().nw_last_error().decode(errors="replace"))

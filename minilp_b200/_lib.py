"""ctypes binding of libminilp_b200.so (include/minilp_b200.h).  No torch types cross this boundary."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libminilp_b200.so")
_lib = None

i32, i64, f64, vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p
pd, pi64, pi32, pu8 = C.POINTER(f64), C.POINTER(i64), C.POINTER(i32), C.POINTER(C.c_uint8)

MLP_OK, MLP_INFEASIBLE, MLP_UNBOUNDED, MLP_SINGULAR, MLP_NONFINITE = 0, 1, 2, 3, 4
MLP_INVALID, MLP_CUDA_ERROR, MLP_NO_DEVICE, MLP_NOMEM = 5, 6, 7, 8
AT_MIN, AT_MAX, BASIC, FIXED = 1, 2, 4, 8


class InitState(C.Structure):
    _fields_ = [("orig_var_mins", pd), ("orig_var_maxs", pd), ("orig_obj_coeffs", pd), ("orig_rhs", pd),
                ("nb_vars", pi64), ("nb_var_vals", pd), ("nb_var_obj_coeffs", pd), ("nb_var_states", pu8),
                ("primal_edge_sq_norms", pd), ("basic_vars", pi64), ("basic_var_vals", pd), ("basic_var_mins", pd),
                ("basic_var_maxs", pd), ("dual_edge_sq_norms", pd), ("enable_primal_steepest_edge", i32),
                ("enable_dual_steepest_edge", i32)]


class Entering(C.Structure):
    _fields_ = [("var", i64), ("pos", i64), ("obj_coeff", f64), ("score", f64), ("cur_val", f64)]


class Leaving(C.Structure):
    _fields_ = [("row", i64), ("coeff", f64), ("leaving_new_val", f64), ("basic_val", f64), ("ties", i64), ("near_ties", i64)]


class DualRow(C.Structure):
    _fields_ = [("row", i64), ("val", f64), ("min", f64), ("max", f64)]


class DualEntering(C.Structure):
    _fields_ = [("var", i64), ("pos", i64), ("coeff", f64), ("obj_coeff", f64), ("cur_val", f64), ("ties", i64),
                ("near_ties", i64)]


class PivotInfo(C.Structure):
    _fields_ = [("entering_var", i64), ("col", i64), ("entering_obj_coeff", f64), ("entering_new_val", f64), ("entering_diff", f64), ("has_elem", i32),
                ("row", i64), ("coeff", f64), ("leaving_new_val", f64), ("refactor", i32)]


class PivotResult(C.Structure):
    _fields_ = [("leaving_var", i64), ("col_nnz", i64), ("eta_count", i64), ("lu_nnz", i64), ("refactored", i32)]


class VarInfo(C.Structure):
    _fields_ = [("flags", C.c_uint32), ("pos_or_row", i64), ("obj_coeff", f64), ("value", f64)]


class AddRowResult(C.Structure):
    _fields_ = [("row", i64), ("slack_var", i64), ("basic_val", f64), ("rhs", f64), ("lu_nnz", i64)]


class Counters(C.Structure):
    _fields_ = [("kernel_launches", i64), ("h2d_bytes", i64), ("d2h_bytes", i64), ("refactors", i64), ("etas_pushed", i64),
                ("k_structural", i64), ("lu_nnz", i64), ("eta_count", i64), ("ratio_ties", i64), ("ratio_near_ties", i64), ("refreshes", i64), ("refresh_rejects", i64)]


class Profile(C.Structure):
    _fields_ = [("price_v_ms", f64), ("price_rho_ms", f64), ("price_v_launches", i64), ("price_rho_launches", i64),
                ("price_v_bytes", i64), ("price_rho_bytes", i64)]


# every symbol include/minilp_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "mlp_last_error": (C.c_char_p, []),
    "mlp_version": (C.c_char_p, []),
    "mlp_device_count": (C.c_int, []),
    "mlp_engine_create_dense": (i32, [C.c_int, i64, i64, C.POINTER(vp)]),
    "mlp_engine_destroy": (None, [vp]),
    "mlp_engine_create_sparse": (i32, [C.c_int, i64, i64, i64, pi64, pi32, pd, C.POINTER(vp)]),
    "mlp_solver_create_sparse": (i32, [C.c_int, i64, i64, i64, pi64, pi32, pd, C.POINTER(vp)]),
    "mlp_engine_download_csc": (i32, [vp, pi64, pi32, pd]),
    "mlp_engine_create_sparse_sharded": (i32, [C.c_int, i64, i64, i64, pi64, pi32, pd, i32, i32, i32, vp, C.POINTER(vp)]),
    "mlp_solver_create_sparse_sharded": (i32, [C.c_int, i64, i64, i64, pi64, pi32, pd, i32, i32, i32, vp, C.POINTER(vp)]),
    "mlp_engine_create_dense_sharded": (i32, [C.c_int, i64, i64, i32, i32, i32, vp, C.POINTER(vp)]),
    "mlp_nccl_get_unique_id": (i32, [vp]),
    "mlp_local_group_create": (i32, [i32, C.POINTER(vp)]),
    "mlp_local_group_destroy": (None, [vp]),
    "mlp_engine_local_range": (i32, [vp, pi64, pi64]),
    "mlp_engine_exchange_kind": (i32, [vp]),
    "mlp_engine_upload_local_rows": (i32, [vp, i64, i64, pd]),
    "mlp_engine_upload_rows": (i32, [vp, i64, i64, pd]),
    "mlp_engine_init_state": (i32, [vp, C.POINTER(InitState)]),
    "mlp_engine_set_primal_steepest_edge": (i32, [vp, i32]),
    "mlp_refactor": (i32, [vp, pi64]),
    "mlp_select_entering_primal": (i32, [vp, C.POINTER(Entering)]),
    "mlp_ftran_col": (i32, [vp, i64]),
    "mlp_ratio_primal": (i32, [vp, i32, f64, C.POINTER(Leaving)]),
    "mlp_btran_unit": (i32, [vp, i64]),
    "mlp_price_row": (i32, [vp]),
    "mlp_calc_row_coeffs": (i32, [vp, i64]),
    "mlp_select_row_dual": (i32, [vp, C.POINTER(DualRow)]),
    "mlp_ratio_dual": (i32, [vp, i64, f64, C.POINTER(DualEntering)]),
    "mlp_dual_select_ratio": (i32, [vp, C.POINTER(DualRow), C.POINTER(DualEntering)]),
    "mlp_pivot": (i32, [vp, C.POINTER(PivotInfo), C.POINTER(PivotResult)]),
    "mlp_recalc_obj_coeffs": (i32, [vp, pd]),
    "mlp_recalc_basic_vals": (i32, [vp]),
    "mlp_solver_set_recalc_period": (None, [vp, i64]),
    "mlp_solver_set_refactor_factor": (None, [vp, f64]),
    "mlp_solver_recalcs_done": (i64, [vp]),
    "mlp_engine_clone": (i32, [vp, C.POINTER(vp)]),
    "mlp_solver_clone": (i32, [vp, C.POINTER(vp)]),
    "mlp_get_var": (i32, [vp, i64, C.POINTER(VarInfo)]),
    "mlp_set_nb_state": (i32, [vp, i64, C.c_uint32]),
    "mlp_engine_add_row": (i32, [vp, pd, pd, f64, f64, f64, C.POINTER(AddRowResult)]),
    "mlp_solver_add_constraint": (i32, [vp, i64, pi64, pd, i32, f64]),
    "mlp_solver_fix_var": (i32, [vp, i64, f64]),
    "mlp_solver_unfix_var": (i32, [vp, i64, pi32]),
    "mlp_solver_add_gomory_cut": (i32, [vp, i64]),
    "mlp_download_f64": (i32, [vp, i32, pd, i64]),
    "mlp_download_basic_vars": (i32, [vp, pi64]),
    "mlp_download_var_state": (i32, [vp, pu8, pi32]),
    "mlp_get_counters": (i32, [vp, C.POINTER(Counters)]),
    "mlp_engine_stream": (vp, [vp]),
    "mlp_engine_sync": (i32, [vp]),
    "mlp_bench_price_dense": (i32, [vp, i32, pd, pi64]),
    "mlp_engine_set_tuning": (i32, [vp, i32, i32]),
    "mlp_engine_get_tuning": (i32, [vp, i32, C.POINTER(i32)]),
    "mlp_event_mark": (i32, [vp, i32]),
    "mlp_event_elapsed_ms": (i32, [vp, i32, i32, pd]),
    "mlp_profile_enable": (i32, [vp, i32]),
    "mlp_profile_get": (i32, [vp, C.POINTER(Profile)]),
    "mlp_solver_create_dense": (i32, [C.c_int, i64, i64, C.POINTER(vp)]),
    "mlp_solver_destroy": (None, [vp]),
    "mlp_solver_create_dense_sharded": (i32, [C.c_int, i64, i64, i32, i32, i32, vp, C.POINTER(vp)]),
    "mlp_solver_upload_local_rows": (i32, [vp, i64, i64, pd]),
    "mlp_solver_engine": (vp, [vp]),
    "mlp_solver_upload_rows": (i32, [vp, i64, i64, pd]),
    "mlp_solver_init": (i32, [vp, pd, pd, pd, pi32, pd]),
    "mlp_solver_run": (i32, [vp, i64, pi32]),
    "mlp_solver_cur_obj_val": (f64, [vp]),
    "mlp_solver_pivots_done": (i64, [vp]),
    "mlp_solver_num_vars": (i64, [vp]),
    "mlp_solver_num_constraints": (i64, [vp]),
    "mlp_solver_values": (i32, [vp, pd]),
    "mlp_solver_trace_len": (i64, [vp]),
    "mlp_solver_get_trace": (i64, [vp, i64, i64, pd]),
    "mlp_solver_set_record_trace": (None, [vp, i32]),
    "mlp_solver_get_nb_vars": (i32, [vp, pi64]),
    "mlp_solver_get_basic_vars": (i32, [vp, pi64]),
    "mlp_solver_timers": (None, [vp, pd, pd]),
    "mlp_solver_tie_stats": (None, [vp, pi64]),
    "mlp_mps_parse": (i32, [C.c_char_p, i64, C.POINTER(vp)]),
    "mlp_mps_free": (None, [vp]),
    "mlp_mps_name": (C.c_char_p, [vp]),
    "mlp_mps_sizes": (None, [vp, pi64, pi64, pi64, pi64]),
    "mlp_mps_export": (i32, [vp, pd, pd, pd, pi64, pi32, pd, pi32, pd, C.c_char_p, pi64]),
    "mlp_shard_range": (None, [i64, i32, i32, pi64, pi64]),
    "mlp_reduce_candidates": (i32, [pd, pi64, pi64, i32]),
    "mlp_synth_rows": (None, [i32, i64, i64, C.c_uint64, i64, i64, i32, pd]),
    "mlp_synth_block": (None, [i32, i64, i64, C.c_uint64, i64, i64, i64, i64, i32, pd]),
    "mlp_synth_vectors": (i32, [i32, i64, i64, C.c_uint64, pd, pd, pd, pi32, pd]),
}


def build(force=False):
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU).  make decides what is stale:
    its prerequisite list names every source, so no second list can fall out of date here."""
    csrc = os.path.join(_HERE, "csrc")
    cmd = ["make", "-C", csrc, "-s"] + (["-B"] if force else [])
    subprocess.check_call(cmd)
    return LIB_PATH

def lib():
    """Load libminilp_b200.so; fails loudly if it is missing (there is no Python/CPU fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(minilp_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L

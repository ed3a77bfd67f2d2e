"""ctypes binding of libporoelas_b200.so (include/poroelas_b200.h).  Thin: every call goes straight
to the C ABI; there is no Python or CPU compute path.  Importing works without a GPU (so that the
symbol table can be checked); creating a handle does not."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libporoelas_b200.so")

PE_BR1_WG, PE_CGQ1_WG, PE_WG_Q2RT2 = 0, 1, 2
PE_CASE_ZERO, PE_CASE_COSCOS, PE_CASE_DARCY_COSCOS = 0, 1, 2
PE_FACE_DIR_U, PE_FACE_DIR_P, PE_FACE_NEU_U = 1, 2, 4
PE_CFG_HOST_SETUP_ONLY = 1
PE_OK, PE_ERR_ARG, PE_ERR_NO_DEVICE, PE_ERR_CUDA, PE_ERR_PATTERN, PE_ERR_NOT_CONVERGED = 0, -1, -2, -3, -4, -5


class pe_config(C.Structure):
    _fields_ = [("element", C.c_int32), ("dim", C.c_int32), ("device", C.c_int32), ("rank", C.c_int32),
                ("world_size", C.c_int32), ("flags", C.c_int32), ("reserved", C.c_int32 * 2)]


class pe_sizes(C.Structure):
    _fields_ = [(k, C.c_int64) for k in
                ("n_cells", "n_vertices", "n_loc", "n_u", "n_pint", "n_pface", "n_dofs", "nnz", "cell_begin",
                 "cell_end", "n_owned_rows", "nnz_owned", "n_ghost_cells", "solver_rows", "solver_nnz")] + [("reserved", C.c_int64 * 1)]


class pe_solve_stats(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("converged", C.c_int32), ("rhs_norm", C.c_double),
                ("residual_norm", C.c_double), ("rel_residual", C.c_double), ("seconds", C.c_double),
                ("spmv_count", C.c_int32), ("restarts", C.c_int32)]


# every symbol include/poroelas_b200.h declares
EXPORTS = [
    "pe_create", "pe_destroy", "pe_last_error", "pe_version", "pe_make_unit_square", "pe_make_unit_cube", "pe_set_mesh3", "pe_get_mesh3", "pe_local_matrices3", "pe_set_boundary3", "pe_set_mesh",
    "pe_distribute_dofs", "pe_set_dofs", "pe_make_sparsity_pattern", "pe_set_pattern", "pe_get_sizes", "pe_get_owned_ranges",
    "pe_get_mesh", "pe_get_dofs", "pe_get_pattern", "pe_get_dealii_block_order", "pe_set_material",
    "pe_set_lognormal_k", "pe_set_case", "pe_set_boundary", "pe_assemble_system", "pe_assemble_rhs_only",
    "pe_get_matrix", "pe_get_rhs", "pe_solve_time_step", "pe_set_solver_options", "pe_set_initial_guess", "pe_set_recycling", "pe_set_krylov", "pe_set_tuning", "pe_get_solution", "pe_get_solution_global", "pe_set_solution", "pe_spmv",
    "pe_bench_spmv", "pe_bench_fp64", "pe_local_matrices", "pe_local_matrices_bench", "pe_step_errors", "pe_postprocess", "pe_darcy_errors", "pe_comm_unique_id", "pe_comm_init",
    "pe_get_timings",
]

_lib = None


class PoroElasError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("poroelas_b200 error %d: %s" % (code, msg))
        self.code = code


def load():
    """Load the CUDA library.  Fails loudly if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libporoelas_b200.so is not built (run `make -C poroelasticity_b200/csrc` or "
                          "__graft_entry__.build()); poroelasticity_b200 has no CPU or Python fallback")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_uint64
    sig = {
        "pe_create": [C.POINTER(vp), C.POINTER(pe_config)],
        "pe_destroy": [vp],
        "pe_make_unit_square": [vp, i32, dbl, u64, i32],
        "pe_set_mesh": [vp, i64, vp, vp, i64, i64, vp],
        "pe_make_unit_cube": [vp, i32, dbl, u64],
        "pe_set_mesh3": [vp, i64, vp, vp, vp, i64, vp],
        "pe_get_mesh3": [vp, vp, vp, vp, vp, vp, vp],
        "pe_local_matrices3": [vp, i64, i64, dbl, vp, vp, vp],
        "pe_set_boundary3": [vp, vp, vp],
        "pe_distribute_dofs": [vp],
        "pe_set_dofs": [vp, i64, i64, i64, vp],
        "pe_make_sparsity_pattern": [vp],
        "pe_set_pattern": [vp, vp, vp],
        "pe_get_sizes": [vp, C.POINTER(pe_sizes)],
        "pe_get_owned_ranges": [vp, vp, vp],
        "pe_get_mesh": [vp, vp, vp, vp, vp],
        "pe_get_dofs": [vp, vp],
        "pe_get_pattern": [vp, vp, vp],
        "pe_get_dealii_block_order": [vp, vp],
        "pe_set_material": [vp, dbl, dbl, dbl, dbl, dbl, vp],
        "pe_set_lognormal_k": [vp, dbl, u64],
        "pe_set_case": [vp, i32, vp],
        "pe_set_boundary": [vp, i64, vp, vp, vp, vp],
        "pe_assemble_system": [vp, dbl, dbl, vp],
        "pe_assemble_rhs_only": [vp, dbl, dbl, vp],
        "pe_get_matrix": [vp, vp],
        "pe_get_rhs": [vp, vp],
        "pe_solve_time_step": [vp, dbl, i32, vp, C.POINTER(pe_solve_stats)],
        "pe_set_solver_options": [vp, i32, dbl, dbl, i32],
        "pe_set_initial_guess": [vp, i32],
        "pe_set_recycling": [vp, i32],
        "pe_set_krylov": [vp, i32, i32],
        "pe_set_tuning": [vp, C.c_char_p, dbl],
        "pe_get_solution": [vp, vp],
        "pe_set_solution": [vp, vp],
        "pe_get_solution_global": [vp, vp],
        "pe_spmv": [vp, vp, vp],
        "pe_bench_spmv": [vp, i32, C.POINTER(dbl)],
        "pe_local_matrices": [vp, i64, i64, dbl, dbl, vp, vp, vp],
        "pe_local_matrices_bench": [vp, dbl, dbl, i64, i32, C.POINTER(dbl), C.POINTER(dbl)],
        "pe_step_errors": [vp, dbl, vp],
        "pe_postprocess": [vp, vp, vp, vp, vp],
        "pe_darcy_errors": [vp, vp],
        "pe_comm_unique_id": [vp],
        "pe_comm_init": [vp, vp],
        "pe_get_timings": [vp, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(dbl), C.POINTER(dbl), C.POINTER(i64)],
    }
    for name, args in sig.items():
        f = getattr(L, name)
        f.argtypes = args
        f.restype = C.c_int
    L.pe_last_error.argtypes = [vp]
    L.pe_last_error.restype = C.c_char_p
    L.pe_version.restype = C.c_char_p
    _lib = L
    return L


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)

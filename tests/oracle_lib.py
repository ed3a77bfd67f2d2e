"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs -- never by
the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_lib = None

BR1_WG, CGQ1_WG = 0, 1
CASE_ZERO, CASE_COSCOS = 0, 1
FK_DIR_U, FK_DIR_P, FK_NEU_U = 1, 2, 4


def build():
    subprocess.check_call(["make", "-s", "-C", _ORACLE_DIR])


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(_ORACLE_DIR, "liboracle.so")
        if not os.path.exists(so):
            build()
        _lib = C.CDLL(so)
        _lib.orc_create.restype = C.c_void_p
        _lib.orc_create.argtypes = [C.c_int, C.c_int, C.c_double, C.c_uint64, C.c_int]
        _lib.orc_destroy.argtypes = [C.c_void_p]
        _lib.orc_set_material.argtypes = [C.c_void_p] + [C.c_double] * 5
        _lib.orc_set_case.argtypes = [C.c_void_p, C.c_int] + [C.c_double] * 5
        _lib.orc_set_kcell.argtypes = [C.c_void_p, C.c_void_p]
        _lib.orc_set_traction.argtypes = [C.c_void_p, C.c_double, C.c_double]
        _lib.orc_set_vertices.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_set_face_kind.argtypes = [C.c_void_p, C.c_void_p]
        _lib.orc_sizes.argtypes = [C.c_void_p, C.c_void_p]
        _lib.orc_get_mesh.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        _lib.orc_get_dofs.argtypes = [C.c_void_p, C.c_void_p]
        _lib.orc_get_pattern.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_local.argtypes = [C.c_void_p, C.c_int, C.c_double] + [C.c_void_p] * 5
        _lib.orc_local_matrices.restype = C.c_double
        _lib.orc_local_matrices.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p]
        _lib.orc_assemble.restype = C.c_double
        _lib.orc_assemble.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_step_errors.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        _lib.orc_postprocess.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        _lib.orc_count_steps.restype = C.c_int
        _lib.orc_count_steps.argtypes = [C.c_double, C.c_double]
        _lib.orc_gauss_jordan.argtypes = [C.c_void_p, C.c_int]
        _lib.orc_qgauss.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        _lib.orc_u01.restype = C.c_double
        _lib.orc_u01.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One reference problem on the unit square (see oracle/oracle_capi.cpp)."""

    def __init__(self, element, n_refine, distort=0.0, seed=12345, bc_mode=0):
        self.L = lib()
        self.h = self.L.orc_create(element, n_refine, distort, seed, bc_mode)
        s = np.zeros(8, np.int64)
        self.L.orc_sizes(self.h, _p(s))
        (self.n_cells, self.n_vertices, self.n_loc, self.n_u, self.n_pint, self.n_pface, self.n_dofs,
         self.nnz) = [int(x) for x in s]
        self.element = element

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def set_material(self, lam, mu, alpha, c0, dt):
        self.L.orc_set_material(self.h, lam, mu, alpha, c0, dt)

    def set_case(self, cid, f_lambda=1.0, f_mu=1.0, f_alpha=1.0, s_capacity=0.0, s_alpha=1.0):
        self.L.orc_set_case(self.h, cid, f_lambda, f_mu, f_alpha, s_capacity, s_alpha)

    def set_kcell(self, k):
        k = np.ascontiguousarray(k, np.float64)
        assert k.size == self.n_cells
        self.L.orc_set_kcell(self.h, _p(k))

    def set_vertices(self, vx, vy):
        vx = np.ascontiguousarray(vx, np.float64)
        vy = np.ascontiguousarray(vy, np.float64)
        assert vx.size == self.n_vertices and vy.size == self.n_vertices
        self.L.orc_set_vertices(self.h, _p(vx), _p(vy))

    def set_traction(self, tx, ty):
        self.L.orc_set_traction(self.h, tx, ty)

    def set_face_kind(self, fk):
        fk = np.ascontiguousarray(fk, np.uint8)
        assert fk.size == 4 * self.n_cells
        self.L.orc_set_face_kind(self.h, _p(fk))

    def mesh(self):
        vx = np.zeros(self.n_vertices)
        vy = np.zeros(self.n_vertices)
        cv = np.zeros((self.n_cells, 4), np.uint32)
        fk = np.zeros((self.n_cells, 4), np.uint8)
        self.L.orc_get_mesh(self.h, _p(vx), _p(vy), _p(cv), _p(fk))
        return vx, vy, cv, fk

    def cell_dofs(self):
        cd = np.zeros((self.n_cells, self.n_loc), np.uint32)
        self.L.orc_get_dofs(self.h, _p(cd))
        return cd

    def pattern(self):
        rp = np.zeros(self.n_dofs + 1, np.int64)
        ci = np.zeros(self.nnz, np.int32)
        self.L.orc_get_pattern(self.h, _p(rp), _p(ci))
        return rp, ci

    def local(self, cell, t, old=None):
        old = np.zeros(self.n_dofs) if old is None else np.ascontiguousarray(old, np.float64)
        A = np.zeros((self.n_loc, self.n_loc))
        b = np.zeros(self.n_loc)
        ic = np.zeros(self.n_loc, np.uint8)
        g = np.zeros(self.n_loc)
        self.L.orc_local(self.h, cell, t, _p(old), _p(A), _p(b), _p(ic), _p(g))
        return A, b, ic.astype(bool), g

    def local_matrices(self, first, n, t=0.0, want=True):
        out = np.zeros((n, self.n_loc, self.n_loc)) if want else None
        sec = self.L.orc_local_matrices(self.h, first, n, t, _p(out) if want else None)
        return out, sec

    def assemble(self, t, old=None):
        old = np.zeros(self.n_dofs) if old is None else np.ascontiguousarray(old, np.float64)
        v = np.zeros(self.nnz)
        r = np.zeros(self.n_dofs)
        sec = self.L.orc_assemble(self.h, t, _p(old), _p(v), _p(r))
        return v, r, sec

    def step_errors(self, t, sol):
        sol = np.ascontiguousarray(sol, np.float64)
        e = np.zeros(3)
        self.L.orc_step_errors(self.h, t, _p(sol), _p(e))
        return e

    def postprocess(self, sol):
        """(div_displacement, displacement_magnitude, darcy beta [n_cells][4], cell-centre velocity [n_cells][2])"""
        sol = np.ascontiguousarray(sol, np.float64)
        d, m = np.zeros(self.n_cells), np.zeros(self.n_cells)
        b, v = np.zeros((self.n_cells, 4)), np.zeros((self.n_cells, 2))
        self.L.orc_postprocess(self.h, _p(sol), _p(d), _p(m), _p(b), _p(v))
        return d, m, b, v

    def csr(self, values):
        import scipy.sparse as sp
        rp, ci = self.pattern()
        return sp.csr_matrix((values, ci, rp), shape=(self.n_dofs, self.n_dofs))


class EsmOracle:
    """EltStiffMat.cc (WGDarcyEquation<2>) on the unit square: oracle/esm_oracle.hpp."""

    def __init__(self, n_refine, distort=0.0, seed=12345):
        self.L = lib()
        L = self.L
        L.orc_esm_create.restype = C.c_void_p
        L.orc_esm_create.argtypes = [C.c_int, C.c_double, C.c_uint64]
        L.orc_esm_destroy.argtypes = [C.c_void_p]
        L.orc_esm_sizes.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_esm_get_mesh.argtypes = [C.c_void_p] * 4
        L.orc_esm_get_dofs.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_esm_get_pattern.argtypes = [C.c_void_p] * 3
        L.orc_esm_set_kcell.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_esm_local_matrices.restype = C.c_double
        L.orc_esm_local_matrices.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_esm_assemble.restype = C.c_double
        L.orc_esm_assemble.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_esm_postprocess.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.h = L.orc_esm_create(n_refine, distort, seed)
        s = np.zeros(5, np.int64)
        L.orc_esm_sizes(self.h, _p(s))
        self.n_cells, self.n_vertices, self.n_loc, self.n_dofs, self.nnz = [int(x) for x in s]

    def __del__(self):
        try:
            self.L.orc_esm_destroy(self.h)
        except Exception:
            pass

    def mesh(self):
        vx, vy = np.zeros(self.n_vertices), np.zeros(self.n_vertices)
        cv = np.zeros((self.n_cells, 4), np.uint32)
        self.L.orc_esm_get_mesh(self.h, _p(vx), _p(vy), _p(cv))
        return vx, vy, cv

    def cell_dofs(self):
        cd = np.zeros((self.n_cells, self.n_loc), np.uint32)
        self.L.orc_esm_get_dofs(self.h, _p(cd))
        return cd

    def pattern(self):
        rp = np.zeros(self.n_dofs + 1, np.int64)
        ci = np.zeros(self.nnz, np.int32)
        self.L.orc_esm_get_pattern(self.h, _p(rp), _p(ci))
        return rp, ci

    def set_kcell(self, k):
        k = np.ascontiguousarray(k, np.float64)
        assert k.size == self.n_cells
        self.L.orc_esm_set_kcell(self.h, _p(k))

    def local_matrices(self, first, n):
        A = np.zeros((n, self.n_loc, self.n_loc))
        b = np.zeros((n, self.n_loc))
        sec = self.L.orc_esm_local_matrices(self.h, first, n, _p(A), _p(b))
        return A, b, sec

    def assemble(self, with_boundary_values=True):
        v, r, s = np.zeros(self.nnz), np.zeros(self.n_dofs), np.zeros(self.n_dofs)
        sec = self.L.orc_esm_assemble(self.h, _p(v), _p(r), _p(s), 1 if with_boundary_values else 0)
        return v, r, s, sec

    def postprocess(self, sol):
        """(L2_error_vel^2, L2_error_flux^2) of EltStiffMat.cc postprocess(), ESM:662-919"""
        sol = np.ascontiguousarray(sol, np.float64)
        e = np.zeros(2)
        self.L.orc_esm_postprocess(self.h, _p(sol), _p(e))
        return e

    def csr(self, values):
        import scipy.sparse as sp
        rp, ci = self.pattern()
        return sp.csr_matrix((values, ci, rp), shape=(self.n_dofs, self.n_dofs))


class Oracle3D:
    """CGQ1-WG cell loop on hexahedra (oracle/poro_oracle_3d.hpp): unit cube, refine_global(n), 31 dofs per cell."""

    n_loc = 31

    def __init__(self, n_refine, distort=0.0, seed=12345):
        self.L = lib()
        L = self.L
        L.orc3_create.restype = C.c_void_p
        L.orc3_create.argtypes = [C.c_int, C.c_double, C.c_uint64]
        L.orc3_destroy.argtypes = [C.c_void_p]
        L.orc3_sizes.argtypes = [C.c_void_p, C.c_void_p]
        L.orc3_get_mesh.argtypes = [C.c_void_p] * 5
        L.orc3_set_material.argtypes = [C.c_void_p] + [C.c_double] * 5
        L.orc3_set_kcell.argtypes = [C.c_void_p, C.c_void_p]
        L.orc3_local.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        self.h = L.orc3_create(n_refine, distort, seed)
        s = np.zeros(2, np.int64)
        L.orc3_sizes(self.h, _p(s))
        self.n_cells, self.n_vertices = int(s[0]), int(s[1])

    def __del__(self):
        try:
            self.L.orc3_destroy(self.h)
        except Exception:
            pass

    def mesh(self):
        x, y, z = np.zeros(self.n_vertices), np.zeros(self.n_vertices), np.zeros(self.n_vertices)
        cv = np.zeros((self.n_cells, 8), np.uint32)
        self.L.orc3_get_mesh(self.h, _p(x), _p(y), _p(z), _p(cv))
        return x, y, z, cv

    def set_material(self, lam, mu, alpha, c0, dt):
        self.L.orc3_set_material(self.h, lam, mu, alpha, c0, dt)

    def set_kcell(self, k):
        k = np.ascontiguousarray(k, np.float64)
        assert k.size == self.n_cells
        self.L.orc3_set_kcell(self.h, _p(k))

    def dofs_pattern(self):
        """(n_u, n_pint, n_pface), cell_dofs [n_cells][31], rowptr, colidx of the 3-D numbering / full-coupling pattern"""
        self.L.orc3_dofs_pattern.argtypes = [C.c_void_p] * 5
        s = np.zeros(4, np.int64)
        self.L.orc3_dofs_pattern(self.h, _p(s), None, None, None)
        cd = np.zeros((self.n_cells, 31), np.uint32)
        rp = np.zeros(int(s[:3].sum()) + 1, np.int64)
        ci = np.zeros(int(s[3]), np.int32)
        self.L.orc3_dofs_pattern(self.h, _p(s), _p(cd), _p(rp), _p(ci))
        return tuple(int(v) for v in s[:3]), cd, rp, ci

    def set_boundary(self, face_kind=None, traction=None):
        """face_kind [n_cells][6] bits FK_DIR_U = 1 | FK_DIR_P = 2 | FK_NEU_U = 4; None = the shipped sandwich boundary"""
        self.L.orc3_set_boundary.argtypes = [C.c_void_p] * 3
        if face_kind is None:
            self.L.orc3_set_boundary(self.h, None, None)
            return
        fk = np.ascontiguousarray(face_kind, np.uint8)
        assert fk.size == 6 * self.n_cells
        tr = np.ascontiguousarray(np.zeros(3) if traction is None else traction, np.float64)
        self.L.orc3_set_boundary(self.h, _p(fk), _p(tr))

    def boundary(self):
        self.L.orc3_get_boundary.argtypes = [C.c_void_p] * 3
        fk, tr = np.zeros((self.n_cells, 6), np.uint8), np.zeros(3)
        self.L.orc3_get_boundary(self.h, _p(fk), _p(tr))
        return fk, tr

    def assemble(self, old_solution=None):
        """(values in CSR order of dofs_pattern(), rhs) of assemble_system in 3-D"""
        self.L.orc3_assemble.argtypes = [C.c_void_p] * 4
        (n_u, n_pi, n_pf), cd, rp, ci = self.dofs_pattern()
        v, r = np.zeros(ci.size), np.zeros(n_u + n_pi + n_pf)
        o = None if old_solution is None else np.ascontiguousarray(old_solution, np.float64)
        self.L.orc3_assemble(self.h, None if o is None else _p(o), _p(v), _p(r))
        return v, r

    def local(self, cell, cell_old=None):
        o = np.zeros(31) if cell_old is None else np.ascontiguousarray(cell_old, np.float64)
        A, b = np.zeros((31, 31)), np.zeros(31)
        self.L.orc3_local(self.h, cell, _p(o), _p(A), _p(b))
        return A, b

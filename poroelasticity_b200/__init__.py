"""poroelasticity_b200: B200 (sm_100a) hot path of sophy1029/PoroElasticity -- Biot assembly + solve.

The product is the C-ABI CUDA library libporoelas_b200.so (include/poroelas_b200.h); this package
is the thin host-side mirror of the reference's class interface used by tests and bench.py.
"""
from . import _capi
from ._capi import (PE_BR1_WG, PE_CGQ1_WG, PE_WG_Q2RT2, PE_CASE_ZERO, PE_CASE_COSCOS, PE_CASE_DARCY_COSCOS,
                    PE_FACE_DIR_U, PE_FACE_DIR_P, PE_FACE_NEU_U, PoroElasError, LIB_PATH, EXPORTS)
from .driver import BiotSystem, PoroElasBR1WG, PoroElas_CGQ1_WG, PoroElas_CGQ1_WG_3D, WGDarcyEquation, comm_unique_id

__all__ = ["BiotSystem", "PoroElasBR1WG", "PoroElas_CGQ1_WG", "PoroElas_CGQ1_WG_3D", "WGDarcyEquation", "PoroElasError", "comm_unique_id"]

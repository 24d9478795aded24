"""ctypes binding of ``libsimudo_b200.so`` (C ABI: ``include/simudo_b200.h``).

The library is the product: there is no CPU fallback.  If the nvcc-built
shared object is missing this module raises at first use, and every plan
insists on CUDA tensors.  (``_use_library_for_tests`` exists only so that the
GPU-less CI container can run the kernel source through ``tests/emul``'s
CUDA-on-CPU shim; nothing in the package calls it.)
"""
import ctypes as C
import os

from .fem.model import SbModel

_HERE = os.path.dirname(os.path.abspath(__file__))
# SIMUDO_B200_SO: developer override (kernel experiments with alternative nvcc builds)
SO_PATH = os.environ.get('SIMUDO_B200_SO') or os.path.join(_HERE, 'csrc', 'libsimudo_b200.so')

_LIB = None
_EMULATED = False

EXPORTED = ['sb_plan_create', 'sb_plan_destroy', 'sb_last_error', 'sb_version',
            'sb_launch_count', 'sb_assemble', 'sb_plan_zero_values', 'sb_plan_nnz',
            'sb_plan_is_native', 'sb_plan_generation_signature', 'sb_locator_create', 'sb_locator_destroy', 'sb_eval_points', 'sb_measure_fp64_peak',
            'sb_plan_column_indices', 'sb_natbc_eval', 'sb_rebase_w', 'sb_newton_update', 'sb_newton_update_damped',
            'sb_surface_flux', 'sb_tables_size', 'sb_plan_set_profiling',
            'sb_plan_last_kernel_ms', 'sb_plan_last_stage_ms', 'sb_band_create', 'sb_band_solve_factored', 'sb_band_destroy', 'sb_band_info',
            'sb_band_solve', 'sb_band_set_refinement', 'sb_csr_residual_dd', 'sb_optics_create',
            'sb_optics_destroy', 'sb_optics_tables_size', 'sb_optics_alpha_rhs',
            'sb_optics_assemble', 'sb_optics_to_cells', 'sb_lcn_iteration']


class SbProblem(C.Structure):
    _fields_ = [
        ('n_cells', C.c_int64), ('n_dofs', C.c_int64), ('P', C.c_int32),
        ('h_cell_vertices', C.c_void_p), ('h_cell_tag', C.c_void_p),
        ('h_cell_dofs', C.c_void_p), ('h_cell_neighbors', C.c_void_p),
        ('h_cell_jump_mask', C.c_void_p), ('h_rowptr', C.c_void_p),
        ('n_patterns', C.c_int64), ('h_patterns', C.c_void_p),
        ('h_cell_pattern', C.c_void_p),
        ('h_tag_params', C.c_void_p),
        ('n_bc', C.c_int64), ('h_bc_dofs', C.c_void_p),
        ('n_natbc', C.c_int64), ('h_natbc_cell', C.c_void_p),
        ('h_natbc_row', C.c_void_p),
        ('h_tables', C.c_void_p), ('n_tables', C.c_int64)]


class SbOpticsProblem(C.Structure):
    _fields_ = [('n_cells', C.c_int64), ('n_dofs', C.c_int64), ('nnz', C.c_int64),
                ('h_cell_dofs', C.c_void_p), ('h_pos', C.c_void_p),
                ('h_rowptr', C.c_void_p), ('h_diag_pos', C.c_void_p),
                ('h_tables', C.c_void_p), ('n_tables', C.c_int64)]


class SbState(C.Structure):
    _fields_ = [('d_u', C.c_void_p), ('d_base_w', C.c_void_p),
                ('d_thmeq_phi', C.c_void_p), ('d_phi_opt', C.c_void_p),
                ('d_bc_values', C.c_void_p), ('d_natbc_values', C.c_void_p)]


def _declare(lib):
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
    lib.sb_plan_create.argtypes = [C.POINTER(SbModel), C.POINTER(SbProblem), C.POINTER(vp)]
    lib.sb_plan_create.restype = C.c_int
    lib.sb_plan_destroy.argtypes = [vp]
    lib.sb_plan_destroy.restype = None
    lib.sb_last_error.restype = C.c_char_p
    lib.sb_version.restype = C.c_int
    lib.sb_launch_count.restype = i64
    lib.sb_tables_size.restype = i64
    lib.sb_plan_nnz.argtypes = [vp]
    lib.sb_plan_nnz.restype = i64
    lib.sb_measure_fp64_peak.argtypes = [C.POINTER(C.c_double)]
    lib.sb_measure_fp64_peak.restype = C.c_int
    lib.sb_plan_is_native.argtypes = [vp]
    lib.sb_plan_is_native.restype = C.c_int
    lib.sb_plan_generation_signature.argtypes = [vp]
    lib.sb_plan_generation_signature.restype = C.c_int
    lib.sb_locator_create.argtypes = [vp, vp, C.c_int32, C.c_int32, vp, C.POINTER(vp)]
    lib.sb_locator_create.restype = C.c_int
    lib.sb_locator_destroy.argtypes = [vp]
    lib.sb_locator_destroy.restype = None
    lib.sb_eval_points.argtypes = [vp, vp, C.c_int32, C.c_int32, i64, vp, vp, vp, vp]
    lib.sb_eval_points.restype = C.c_int
    lib.sb_assemble.argtypes = [vp, C.POINTER(SbState), vp, vp, vp]
    lib.sb_assemble.restype = C.c_int
    lib.sb_plan_zero_values.argtypes = [vp, vp, vp]
    lib.sb_plan_zero_values.restype = C.c_int
    lib.sb_plan_column_indices.argtypes = [vp, vp, vp]
    lib.sb_plan_column_indices.restype = C.c_int
    lib.sb_natbc_eval.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, i64, vp, vp]
    lib.sb_natbc_eval.restype = C.c_int
    lib.sb_rebase_w.argtypes = [vp, vp, vp, i32, i64, vp, vp, vp]
    lib.sb_rebase_w.restype = C.c_int
    lib.sb_newton_update.argtypes = [vp, vp, vp, vp, vp, dbl, dbl, dbl, dbl, vp, vp]
    lib.sb_newton_update.restype = C.c_int
    lib.sb_newton_update_damped.argtypes = [vp, vp, vp, vp, vp, dbl, dbl, dbl, dbl, dbl, vp, vp]
    lib.sb_newton_update_damped.restype = C.c_int
    lib.sb_surface_flux.argtypes = [vp, vp, i32, i64, vp, vp, vp, vp, vp]
    lib.sb_surface_flux.restype = C.c_int
    lib.sb_plan_set_profiling.argtypes = [vp, C.c_int]
    lib.sb_plan_set_profiling.restype = C.c_int
    lib.sb_plan_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_double)]
    lib.sb_plan_last_kernel_ms.restype = C.c_int
    lib.sb_plan_last_stage_ms.argtypes = [vp, C.POINTER(C.c_double)]
    lib.sb_plan_last_stage_ms.restype = C.c_int
    lib.sb_band_create.argtypes = [i64, vp, vp, vp, i32, i64, C.POINTER(vp)]
    lib.sb_band_create.restype = C.c_int
    lib.sb_band_destroy.argtypes = [vp]
    lib.sb_band_destroy.restype = None
    lib.sb_band_info.argtypes = [vp, C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)]
    lib.sb_band_info.restype = C.c_int
    lib.sb_band_solve.argtypes = [vp, i32, vp, i64, vp, vp, vp, vp]
    lib.sb_band_solve.restype = C.c_int
    lib.sb_band_solve_factored.argtypes = [vp, C.c_int32, vp, vp, vp]
    lib.sb_band_solve_factored.restype = C.c_int
    lib.sb_csr_residual_dd.argtypes = [i64, vp, vp, vp, vp, vp, vp, vp]
    lib.sb_csr_residual_dd.restype = C.c_int
    lib.sb_band_set_refinement.argtypes = [vp, i32]
    lib.sb_band_set_refinement.restype = C.c_int
    lib.sb_optics_create.argtypes = [vp, C.POINTER(SbOpticsProblem), C.POINTER(vp)]
    lib.sb_optics_create.restype = C.c_int
    lib.sb_optics_destroy.argtypes = [vp]
    lib.sb_optics_destroy.restype = None
    lib.sb_optics_tables_size.restype = i64
    lib.sb_optics_alpha_rhs.argtypes = [vp, C.POINTER(SbState), i32, vp, vp]
    lib.sb_optics_alpha_rhs.restype = C.c_int
    lib.sb_optics_assemble.argtypes = [vp, vp, dbl, dbl, i64, vp, vp, vp, vp, vp]
    lib.sb_optics_assemble.restype = C.c_int
    lib.sb_optics_to_cells.argtypes = [vp, vp, vp, vp]
    lib.sb_optics_to_cells.restype = C.c_int
    lib.sb_lcn_iteration.argtypes = [C.POINTER(SbModel), i64, vp, vp, vp, vp, vp, dbl, dbl, dbl,
                                     dbl, vp, vp, vp, vp]
    lib.sb_lcn_iteration.restype = C.c_int
    return lib


def get():
    """Load the CUDA extension or fail loudly."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                'simudo_b200 CUDA extension not built: {} is missing. Run '
                '`python -m simudo_b200.csrc.build` (needs nvcc); there is no '
                'CPU fallback for the assembly path.'.format(SO_PATH))
        _LIB = _declare(C.CDLL(SO_PATH))
    return _LIB


def is_emulated():
    return _EMULATED


def check(rc):
    """Map a negative status to RuntimeError, as dolfin/PETSc failures surface
    in the reference (``simudo/fem/adaptive_stepper.py:212-217``)."""
    if rc != 0:
        msg = get().sb_last_error()
        raise RuntimeError('simudo_b200: {} (status {})'.format(
            msg.decode() if msg else 'error', rc))


def _use_library_for_tests(path):
    """TEST ONLY: bind an alternative build of the same C ABI (tests/emul)."""
    global _LIB, _EMULATED
    _LIB = _declare(C.CDLL(path))
    _EMULATED = True
    return _LIB


def _reset_for_tests():
    global _LIB, _EMULATED
    _LIB = None
    _EMULATED = False

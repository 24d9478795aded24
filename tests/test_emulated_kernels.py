"""Kernel *source* exercised on the CPU through tests/emul's CUDA shim.

This is not the product path (the product only loads the nvcc-built library
and demands a CUDA device); it gives the GPU-less container line coverage of
csrc/pdd_kernels.cu + pdd_device.cuh against the oracle.  The real parity
gate is tests/test_gpu_parity.py on the B200.
"""
import pytest

import hotpath_cases as hc
from simudo_b200 import synthetic as syn


@pytest.mark.parametrize('maker,nx,ny,numbering,bcs,pattern', [
    (syn.pn_problem, 5, 3, 'b200', True, 'structural'),
    (syn.pn_problem, 4, 2, 'ufc', True, 'dolfin'),
    (syn.ib_problem, 6, 3, 'b200', True, 'dolfin'),
    (syn.ib_problem, 5, 2, 'ufc', False, 'structural'),
    (syn.pn_problem, 34, 3, 'b200', True, 'structural'),   # > 1 CTA, ragged last CTA
    # closed-form kernels (csrc/native.inl): CSR 'native' and block CSR 'bsr3'
    (syn.ib_problem, 6, 3, 'b200', True, 'native'),
    (syn.ib_problem, 6, 3, 'b200', True, 'bsr3'),
    (syn.pn_problem, 36, 4, 'b200', True, 'bsr3'),         # several tiles, ragged last tile
    (syn.ib_problem, 5, 2, 'b200', False, 'bsr3'),
])
def test_assembly_parity_emulated(emulated_lib, maker, nx, ny, numbering, bcs, pattern):
    hc.assembly_parity(maker, nx, ny, numbering=numbering, with_bcs=bcs, pattern=pattern)


@pytest.mark.parametrize('pattern', ['native', 'bsr3'])
def test_interior_facet_dirichlet_dofs_on_the_closed_form_path_emulated(emulated_lib, pattern):
    """Dirichlet dofs on INTERIOR facets (IB/j = 0 on the bounds of the IB layer,
    fourlayer.py's F.IB_bounds): both cells of such a facet are boundary cells; the shared
    (facet, facet) block carries the identity entries of both (diagonal 2) and the
    plan still runs the closed-form kernels."""
    pa, st, plan, vals, b, vals0, b0 = hc.assembly_parity(
        syn.ib_problem, 9, 4, pattern=pattern, ib_bounds_bc=True)
    assert plan.is_native()
    import numpy as np
    interior_bc = [d for d in pa.bc_dofs if d >= 6 * pa.P * pa.n_cells]
    fc = pa.mesh.facet_cells
    nf = pa.mesh.num_facets
    F0 = 6 * pa.P * pa.n_cells
    on_interior = [d for d in interior_bc if fc[((d - F0) % (3 * nf)) // 3, 1] >= 0]
    assert len(on_interior) >= 6
    col, row = pa.csr.column_indices(want_rows=True)
    for d in on_interior[:6]:
        diag = vals[(row == d) & (col == d)]
        assert len(diag) == 1 and diag[0] == 2.0
        assert np.abs(vals[(row == d) & (col != d)]).max() == 0.0


def test_generation_signature_emulated(emulated_lib):
    hc.generation_signature_matches_interpreter()


def test_thermal_equilibrium_mode_emulated(emulated_lib):
    hc.thermal_equilibrium_mode()


def test_fill_dependent_ib_mobility_emulated(emulated_lib):
    hc.fill_dependent_ib_mobility()


def test_unstaged_rank_tables_emulated(emulated_lib):
    hc.unstaged_rank_tables_match_staged()


def test_non_default_quadrature_emulated(emulated_lib):
    hc.non_default_quadrature_degrees()


def test_deterministic_emulated(emulated_lib):
    hc.assembly_is_deterministic_and_idempotent()


def test_residual_only_emulated(emulated_lib):
    hc.residual_only_and_matrix_only()


def test_rebase_emulated(emulated_lib):
    hc.rebase_w_parity()


def test_newton_update_emulated(emulated_lib):
    hc.newton_update_parity()


def test_natbc_eval_emulated(emulated_lib):
    hc.natbc_eval_parity()


def test_eval_points_emulated(emulated_lib):
    hc.eval_points_parity()


def test_surface_flux_emulated(emulated_lib):
    hc.surface_flux_parity()


def test_band_solver_with_refinement_emulated(emulated_lib):
    """Banded LU + double-double refinement vs scipy on a Newton matrix."""
    import numpy as np
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from parity_util import run_plan
    from simudo_b200.fem.assembly import AssemblyPlan
    from simudo_b200.fem.linear_solver import BandSolver
    pa = syn.pn_problem(nx=12, ny=2)
    st = syn.synthetic_state(pa)
    plan = AssemblyPlan(pa)
    vals, b = run_plan(plan, st)
    bs = BandSolver(pa, 'cpu')
    assert bs.kl < 60 and bs.ku < 60
    x = bs.solve(vals, b).numpy()
    A = sp.csr_matrix((vals.numpy(), pa.csr.column_indices(), pa.csr.rowptr),
                      shape=(pa.N, pa.N))
    r = A @ x - b.numpy()
    rs = 1.0 / np.abs(A).max(axis=1).toarray().ravel()
    assert np.abs(rs * r).max() < 1e-12 * np.abs(rs * b.numpy()).max() + 1e-20


def test_band_window_kernels_emulated(emulated_lib):
    hc.band_window_kernels_match_global_kernels()


def test_optics_parity_emulated(emulated_lib):
    hc.optics_parity(nx=7, ny=3)


def test_lcn_iteration_parity_emulated(emulated_lib):
    hc.lcn_iteration_parity()


def test_bulk_store_matches_entry_flush_emulated(emulated_lib):
    hc.bulk_store_matches_entry_flush()


def test_band_solve_factored_needs_a_factorisation(emulated_lib):
    """sb_band_solve_factored before any sb_band_solve is an argument error, not garbage."""
    import numpy as np
    import torch
    from simudo_b200.fem.linear_solver import BandSolver
    n = 6
    rowptr = np.arange(n + 1, dtype=np.int64)
    bs = BandSolver(None, 'cpu', csr=(n, rowptr, np.arange(n, dtype=np.int32),
                                      np.arange(n, dtype=np.int32)))
    with pytest.raises(RuntimeError):
        bs.solve_factored(torch.ones(n, dtype=torch.float64))
    x = bs.solve(torch.full((n,), 2.0, dtype=torch.float64), torch.ones(n, dtype=torch.float64))
    assert np.allclose(x.numpy(), 0.5)
    x2 = bs.solve_factored(torch.full((n,), 3.0, dtype=torch.float64))
    assert np.allclose(x2.numpy(), 1.5)

"""Parity tests proper: the CUDA path on a B200, through the C ABI, against the
oracle (bit-exact pattern / indices, 1e-12 row-relative values)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', autouse=True)
def real_library():
    from simudo_b200 import _lib
    _lib._reset_for_tests()
    lib = _lib.get()                       # raises if the .so is missing
    assert not _lib.is_emulated()
    yield lib


import hotpath_cases as hc          # noqa: E402
from simudo_b200 import synthetic as syn   # noqa: E402


@pytest.mark.parametrize('maker,nx,ny,numbering,bcs,pattern', [
    (syn.pn_problem, 5, 3, 'b200', True, 'structural'),
    (syn.pn_problem, 4, 2, 'ufc', True, 'dolfin'),
    (syn.pn_problem, 67, 2, 'b200', True, 'dolfin'),       # config-1 size (132 cells)
    (syn.pn_problem, 67, 2, 'ufc', True, 'structural'),
    (syn.ib_problem, 6, 3, 'b200', True, 'dolfin'),
    (syn.ib_problem, 5, 2, 'ufc', False, 'structural'),
    (syn.ib_problem, 279, 2, 'b200', True, 'structural'),  # config-2 size (556 cells)
    (syn.ib_problem, 279, 2, 'ufc', True, 'dolfin'),
    (syn.pn_problem, 40, 33, 'b200', True, 'structural'),  # 2 496 cells, many CTAs
    (syn.ib_problem, 30, 21, 'b200', True, 'structural'),
    (syn.ib_problem, 30, 21, 'b200', True, 'dolfin'),
    # closed-form kernels (csrc/native.inl)
    (syn.pn_problem, 5, 3, 'b200', True, 'native'),
    (syn.ib_problem, 6, 3, 'b200', True, 'native'),
    (syn.ib_problem, 5, 2, 'b200', False, 'native'),
    (syn.pn_problem, 67, 2, 'b200', True, 'native'),
    (syn.ib_problem, 279, 2, 'b200', True, 'native'),
    (syn.pn_problem, 40, 33, 'b200', True, 'native'),
    (syn.ib_problem, 30, 21, 'b200', True, 'native'),
    (syn.ib_problem, 30, 21, 'ufc', True, 'native'),       # 'native' columns, generic kernels
    # block CSR with 3 x 3 blocks, closed-form kernels, every run leaves by cp.async.bulk
    (syn.pn_problem, 5, 3, 'b200', True, 'bsr3'),
    (syn.ib_problem, 6, 3, 'b200', True, 'bsr3'),
    (syn.ib_problem, 5, 2, 'b200', False, 'bsr3'),
    (syn.pn_problem, 67, 2, 'b200', True, 'bsr3'),
    (syn.ib_problem, 279, 2, 'b200', True, 'bsr3'),
    (syn.pn_problem, 40, 33, 'b200', True, 'bsr3'),
    (syn.ib_problem, 30, 21, 'b200', True, 'bsr3'),
])
def test_assembly_parity(maker, nx, ny, numbering, bcs, pattern):
    hc.assembly_parity(maker, nx, ny, numbering=numbering, with_bcs=bcs, pattern=pattern)


def test_assembly_parity_at_bench_scale():
    """bench.py's configuration (IB cell, 708 x 708 points, 1 003 940 cells, 1.07e9 non-zeros)
    against the oracle; SIMUDO_B200_SCALE_N=<points per axis> shrinks it for quick runs."""
    import os
    n = int(os.environ.get('SIMUDO_B200_SCALE_N', '708'))
    cells, ea, ee, eb = hc.assembly_parity_at_scale(n, n)
    print('at scale: %d cells, row-relative %.2e, entry-relative %.2e, residual %.2e'
          % (cells, ea, ee, eb))
    assert cells > 2 * (n - 1) ** 2 - 1


@pytest.mark.parametrize('pattern', ['native', 'bsr3'])
def test_interior_facet_dirichlet_dofs_on_the_closed_form_path(pattern):
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


def test_generation_signature():
    hc.generation_signature_matches_interpreter()


def test_thermal_equilibrium_mode():
    hc.thermal_equilibrium_mode()


def test_fill_dependent_ib_mobility():
    hc.fill_dependent_ib_mobility()


def test_unstaged_rank_tables():
    hc.unstaged_rank_tables_match_staged()


def test_non_default_quadrature():
    hc.non_default_quadrature_degrees()


def test_deterministic():
    hc.assembly_is_deterministic_and_idempotent()


def test_residual_only():
    hc.residual_only_and_matrix_only()


def test_rebase():
    hc.rebase_w_parity()


def test_newton_update():
    hc.newton_update_parity()


def test_natbc_eval():
    hc.natbc_eval_parity()


def test_band_window_kernels():
    hc.band_window_kernels_match_global_kernels()


def test_eval_points():
    hc.eval_points_parity()


def test_surface_flux():
    hc.surface_flux_parity()


@pytest.mark.parametrize('maker,nx,ny', [
    (syn.pn_problem, 224, 224),          # BASELINE config 3: ~100 k cells
    (syn.ib_problem, 160, 160),
])
def test_jacobian_consistent_with_residual_at_scale(maker, nx, ny):
    n = hc.jacobian_consistent_with_residual_large(maker, nx, ny)
    assert n > 40000


def test_misaligned_values_are_refused():
    """The bulk-copy paths need a 16-byte aligned value array: an odd element offset into a
    larger buffer is refused instead of faulting on the device."""
    import torch
    from simudo_b200.fem.assembly import AssemblyPlan
    for pattern in ('bsr3', 'structural'):
        pa = syn.pn_problem(nx=5, ny=3, pattern=pattern)
        st = syn.synthetic_state(pa)
        plan = AssemblyPlan(pa)
        T = plan.tensor
        buf = torch.zeros(plan.nnz + 1, dtype=torch.float64, device=plan.device)
        with pytest.raises(RuntimeError, match='aligned'):
            plan.assemble(T(st['u']), T(st['base_w']), T(st['thmeq_phi']), None, T(st['bc_values']),
                          T(plan.natbc_device_order(st['natbc_values'])), buf[1:], plan.new_vector())


def test_no_cpu_fallback():
    """The product must refuse to run without CUDA tensors."""
    import torch
    from simudo_b200.fem.assembly import AssemblyPlan
    pa = syn.pn_problem(nx=4, ny=2)
    with pytest.raises(RuntimeError):
        AssemblyPlan(pa, device='cpu')


def test_optics_parity():
    hc.optics_parity(nx=33, ny=5)


def test_lcn_iteration_parity():
    hc.lcn_iteration_parity()


def test_bulk_store_matches_entry_flush():
    hc.bulk_store_matches_entry_flush()


def test_pipelined_host_buffer_assembly_matches_direct():
    """PipelinedAssembler (H2D / kernels / D2H on three streams, double buffered):
    every step's residual equals the direct call, also when inputs change per step."""
    import torch
    from simudo_b200.fem.assembly import AssemblyPlan
    from simudo_b200.fem.pipeline import PipelinedAssembler
    from parity_util import run_plan
    pa = syn.ib_problem(nx=40, ny=20)
    st = syn.synthetic_state(pa)
    plan = AssemblyPlan(pa)
    T = plan.tensor
    host = {k: torch.as_tensor(np.ascontiguousarray(st[k])).pin_memory()
            for k in ('u', 'base_w', 'thmeq_phi', 'phi_opt')}
    pipe = PipelinedAssembler(plan, host, T(st['bc_values']),
                              T(plan.natbc_device_order(st['natbc_values'])))
    outs, refs = [], []
    for i in range(5):
        hi = dict(host)
        hi['u'] = (host['u'] * (1.0 + 1e-3 * i)).pin_memory()
        hb = torch.empty(pa.N, dtype=torch.float64).pin_memory()
        pipe.submit(hi, hb)
        outs.append((hi, hb))
    pipe.wait()
    for hi, hb in outs:
        st_i = dict(st)
        st_i['u'] = hi['u'].numpy()
        _, b = run_plan(plan, st_i)
        assert torch.equal(b.cpu(), hb)

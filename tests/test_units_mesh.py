"""Host logic: units, mesh generation, topology algebra, dof maps."""
import numpy as np
import pytest

from simudo_b200.util import make_unit_registry, NameDict
from simudo_b200.mesh import (Product2DMesh, CellRegions, FacetRegions,
                              ConstructionHelperLayeredStructure, Interval,
                              CInterval, Interval1DTag)
from simudo_b200.fem.dofmap import build_cell_dofs, CsrStructure


def test_units_constants_match_reference_files():
    U = make_unit_registry(("mesh_unit = 1 micrometer",))
    # simudo/util/pint/default_en.txt:88, constants_en.txt:8,15,34
    assert U('1 elementary_charge').m_as('C') == pytest.approx(1.602176487e-19, rel=1e-15)
    assert U('1 planck_constant').m_as('J s') == pytest.approx(6.62606957e-34, rel=1e-15)
    assert (U('300 K') * U.boltzmann_constant).m_as('eV') == pytest.approx(
        1.3806488e-23 * 300 / 1.602176487e-19, rel=1e-15)
    eps0 = U('1 vacuum_permittivity').m_as('F/m')
    assert eps0 == pytest.approx(8.8541878e-12, rel=1e-7)
    assert U('1400 cm^2/V/s').m_as('mesh_unit^2/V/s') == pytest.approx(1.4e11)
    assert (1e18 * U('elementary_charge/cm^3')).m_as('elementary_charge/mesh_unit^3') == pytest.approx(1e6)


def test_units_form_factors():
    U = make_unit_registry(("mesh_unit = 1 micrometer",))
    e = 1.602176487e-19
    dd = (U.mesh_unit ** 3 * U('1/mesh_unit^2/eV') * U('A/mesh_unit^2')
          / U('mesh_unit^2/V/s') / U('1/mesh_unit^3'))
    assert dd.m_as("dimensionless") == pytest.approx(1 / e, rel=1e-14)
    c2 = (U.mesh_unit ** 3 * U('A/mesh_unit^2') * U('1/mesh_unit^2/eV') * U('1 V*s*cm'))
    assert c2.m_as('dimensionless') == pytest.approx(1e4 / e, rel=1e-14)
    c1 = (U.mesh_unit ** 3 * U.eV * U('1/A') * U('1 A*mesh_unit^-3/(electron_volt)'))
    assert c1.m_as('dimensionless') == pytest.approx(1.0, rel=1e-14)


def test_units_arithmetic_and_errors():
    U = make_unit_registry(("mesh_unit = 1 micrometer",))
    q = U('1.12 eV') - U('0.12 eV')
    assert q.m_as('eV') == pytest.approx(1.0)
    assert (U.V * 3.0).to('mV').magnitude == pytest.approx(3000.0)
    with pytest.raises(ValueError):
        U('1 eV') + U('1 cm')


def test_namedict_insertion_order():
    class O(object):
        def __init__(self, n): self.name = n
    d = NameDict([O('CB'), O('IB'), O('VB')])
    assert [x.name for x in d] == ['CB', 'IB', 'VB']
    with pytest.raises(KeyError):
        d.add(O('CB'))


def test_product2d_cell_layout_reference_known_answer():
    # simudo/mesh/product2d.py:116-122
    o = Product2DMesh([0, 1, 2, 3], [0, 1, 2])
    assert list(o.cells_at_ix(1)) == [4, 5, 6, 7]
    m = o.mesh
    assert (np.diff(m.cells, axis=1) > 0).all()          # vertex-sorted
    assert m.num_facets == 3 * 2 + 4 * 2 + 6 + 6 - 3     # Euler: V - E + F = 1
    assert m.num_vertices - m.num_facets + m.num_cells == 1


def test_topology_algebra_identities():
    # the set identities exercised by simudo/test/test_topology.py
    ls = ConstructionHelperLayeredStructure()
    U = make_unit_registry(("mesh_unit = 1 micrometer",))
    ls.params = dict(edge_length=0.05, extra_regions=[], mesh_unit=U.mesh_unit, layers=[
        dict(name='a', material='m1', thickness=0.2),
        dict(name='b', material='m2', thickness=0.3),
        dict(name='c', material='m1', thickness=0.1)])
    ls.run()
    md = ls.mesh_data
    R = CellRegions()
    ev = md.evaluate_topology
    assert ev(R.a | R.b | R.c) == ev(R.domain)
    assert ev(R.domain - R.b) == ev(R.a | R.c)
    assert ev(R.a & R.b) == set()
    assert ev(R.a ^ R.domain) == ev(R.b | R.c)
    ab = ev(R.a.boundary(R.b))
    assert ab == ev(R.b.boundary(R.a).flip())
    assert ev(R.a.boundary(R.b).both()) == ab | ev(R.a.boundary(R.b).flip())
    assert {s for _, s in ev(R.a.boundary(R.b).unsigned())} == {1}
    assert ev(R.a.boundary(R.c)) == set()
    # internal facets of the union contain the a|b interface
    assert {fv for fv, _ in ab} <= {fv for fv, _ in ev((R.a | R.b).internal_facets())}
    # materials map to unions of layers
    assert ev(md.material_to_region_map['m1']) == ev(R.a | R.c)


def test_interval_tagging_matches_reference_example_sizes():
    U = make_unit_registry(("mesh_unit = 1 micrometer",))
    length = 0.5
    g = dict(type="geometric", start=0.001, factor=1.2)
    ls = ConstructionHelperLayeredStructure()
    ls.params = dict(edge_length=length / 20, extra_regions=[], mesh_unit=U.mesh_unit, layers=[
        dict(name='pSi', material='Si', thickness=length / 2, mesh=g),
        dict(name='nSi', material='Si', thickness=length / 2, mesh=g)])
    ls.run()
    xs = np.array(ls.interval_1d_tag.coordinates['coordinates'])
    assert len(xs) == 67 and ls.mesh_data.mesh.num_cells == 132
    assert xs[0] == 0 and xs[-1] == pytest.approx(0.5)
    assert np.diff(xs).min() >= 1e-10
    assert np.diff(xs).max() <= length / 20 + 1e-12
    assert xs[1] == pytest.approx(0.001)
    # contacts: one exterior facet each on a nY = 2 product mesh
    R, F = CellRegions(), FacetRegions()
    lc = R.exterior_left.boundary(R.domain)
    ev = ls.mesh_data.evaluate_topology
    (fv, sign), = ev(lc)
    assert (ls.mesh_data.facet_function == fv).sum() == 1 and sign == -1


@pytest.mark.parametrize('numbering', ['ufc', 'b200'])
def test_csr_structure_is_dolfin_dense_cell_pattern(numbering):
    import scipy.sparse as sp
    m = Product2DMesh(np.linspace(0, 1, 5), np.linspace(0, 1, 4)).mesh
    P = 3
    cd, N = build_cell_dofs(m, P, numbering)
    assert sorted(np.unique(cd)) == list(range(N))
    S = CsrStructure(m, cd, P, N, 'dolfin')
    L = 15 * P
    rows = np.repeat(cd, L, axis=1).ravel()
    cols = np.tile(cd, (1, L)).ravel()
    A = sp.coo_matrix((np.ones(rows.size), (rows, cols)), shape=(N, N)).tocsr()
    A.sort_indices()
    assert (A.indptr == S.rowptr).all()
    assert (A.indices == S.column_indices()).all()
    # nnz = L^2 Nc - (3P)^2 Nf_int
    nfi = int((m.facet_cells[:, 1] >= 0).sum())
    assert S.nnz == L * L * m.num_cells - (3 * P) ** 2 * nfi


def test_ufc_numbering_layout():
    m = Product2DMesh(np.linspace(0, 1, 3), np.linspace(0, 1, 3)).mesh
    cd, N = build_cell_dofs(m, 2, 'ufc')
    nc, nf = m.num_cells, m.num_facets
    assert (cd[:, 0] == 3 * np.arange(nc)).all()                 # DG1 of sub-problem 0
    assert (cd[:, 3] == 3 * nc + 3 * m.cell_facets[:, 0]).all()   # BDM facet 0 dof 0
    assert (cd[:, 12] == 3 * nc + 3 * nf + 3 * np.arange(nc)).all()
    assert (cd[:, 15] == 6 * nc + 3 * nf + 3 * np.arange(nc)).all()


@pytest.mark.parametrize('numbering', ['ufc', 'b200'])
def test_structural_pattern_is_the_masked_dolfin_pattern(numbering):
    import scipy.sparse as sp
    from simudo_b200.fem.dofmap import structural_mask
    m = Product2DMesh(np.linspace(0, 1, 5), np.linspace(0, 1, 4)).mesh
    P = 4
    cd, N = build_cell_dofs(m, P, numbering)
    S = CsrStructure(m, cd, P, N, 'structural')
    mask = structural_mask(P)
    # per-cell structural entries (1116 for an IB cell)
    assert mask.sum() == 3 * (12 + 3 + 9) + 12 * 15 + 3 * (3 * (12 + 3 + 9) + 12 * 18)
    ii, jj = np.nonzero(mask)
    A = sp.coo_matrix((np.ones(cd.shape[0] * len(ii)), (cd[:, ii].ravel(), cd[:, jj].ravel())),
                      shape=(N, N)).tocsr()
    A.sort_indices()
    assert (A.indptr == S.rowptr).all()
    assert (A.indices == S.column_indices()).all()
    assert S.patterns.shape[0] <= 16

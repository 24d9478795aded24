"""Flood-fill partitioning of the quasi-Fermi offsets (simudo/fem/offset_partitioning.py:14-169,
the algorithm of its docstring :40-69) -- host logic, no device."""
import numpy as np

from simudo_b200.fem.offset_partitioning import partition_offsets


def _chain(n):
    adj = np.full((n, 3), -1, dtype=np.int64)
    adj[1:, 0] = np.arange(n - 1)
    adj[:-1, 1] = np.arange(1, n)
    return adj


def test_zero_thresholds_are_the_identity_with_bc_override():
    v = np.array([1.0, 2.0, 3.0, 4.0])
    out = partition_offsets(v, _chain(4), (0.0, 0.0), ([3], [9.0]))
    assert (out == [1.0, 2.0, 3.0, 9.0]).all()


def test_close_neighbours_share_the_seed_value_and_jumps_separate_regions():
    v = np.array([100.0, 101.0, 102.0, 101.0, 50.0, 50.5, 10.0, 11.0, 12.0, 11.0])
    out = partition_offsets(v, _chain(10), (2.0, 0.0))
    # the walk starts at cell 0: everything within 2 of the region's marker joins it
    assert (out[:4] == 100.0).all()
    assert (out[4:6] == 50.0).all()
    assert (out[6:] == 10.0).all()
    # idempotent on its own (piecewise constant) output
    assert (partition_offsets(out, _chain(10), (2.0, 0.0)) == out).all()


def test_relative_threshold_and_isolated_cells_keep_their_value():
    v = np.array([1.0, 1.05, 5.0, 100.0, 104.0])
    out = partition_offsets(v, _chain(5), (0.0, 0.06))
    assert out[0] == out[1] == 1.0
    assert out[2] == 5.0                              # no close neighbour: unmarked -> own value
    assert out[3] == out[4] == 100.0


def test_boundary_cells_seed_their_regions_with_the_bc_value():
    v = np.array([0.30, 0.31, 0.29, 0.90, 0.91])
    out = partition_offsets(v, _chain(5), (0.05, 0.0), ([0, 4], [0.28, 1.0]))
    # cells within the threshold of the *marker* (the BC value) join the BC region ...
    assert (out[:3] == 0.28).all()
    assert out[4] == 1.0
    assert out[3] == 0.90                             # ... 0.90 is not within 0.05 of 1.0
    # a BC value further away than the threshold stays alone; its neighbours form their own region
    out1 = partition_offsets(v, _chain(5), (0.05, 0.0), ([0], [0.20]))
    assert out1[0] == 0.20 and out1[1] == out1[2] == 0.31
    out0 = partition_offsets(v, _chain(5), (0.05, 0.0), ([0, 4], [0.25, 1.0]),
                             fill_with_zero_except_bc=True)
    assert (out0 == [0.25, 0.0, 0.0, 0.0, 1.0]).all()


def test_two_dimensional_adjacency_fills_across_both_directions():
    from simudo_b200.mesh import Product2DMesh
    mesh = Product2DMesh(np.linspace(0, 1, 6), np.linspace(0, 1, 5)).mesh
    adj = mesh.cell_to_cells_adjacency_via_facet()
    xc = mesh.coordinates[mesh.cells].mean(axis=1)[:, 0]
    v = np.where(xc > 0.5, 1.0, 0.3) + 1e-4 * np.sin(37.0 * np.arange(mesh.num_cells))
    out = partition_offsets(v, adj, (1e-2, 0.0))
    assert len(np.unique(out)) == 2
    assert len(np.unique(out[xc > 0.5])) == 1 and len(np.unique(out[xc < 0.5])) == 1

#!/usr/bin/env python
"""Dump true dolfin goldens for the assembly path.  NOT EXECUTED in the build
environment (FEniCS absent); run inside the reference's docker image
(``ecdee/simudo``, /root/reference/doc/source/user/install.rst:61-63)::

    python tools/dump_dolfin_reference.py out.npz

It runs the reference's tutorial pn diode up to the 'full' goal, assembles the
Jacobian and residual exactly as ``NewtonSolution.A`` / ``.b`` do
(simudo/fem/newton_solver.py:61-77) at the thermal-equilibrium state with
V_ext = 0.05 V, and stores: mesh coordinates/cells, the mixed dofmap
(cell_dofs), CSR (indptr, indices, data), b, and the solution vector, so that
simudo_b200 can be compared after applying the dof permutation between dolfin's
reordered numbering and the 'ufc' numbering.
"""
import sys

import numpy as np


def main(out):
    import dolfin
    from simudo.example.pn_diode import pn_diode          # noqa: F401 (reference)
    from simudo.fem import NewtonSolver
    loc = pn_diode.run()                                   # returns locals()
    problem = loc['problem']
    pdd = problem.pdd
    loc['V_ext'].magnitude.assign(0.05)
    solver = NewtonSolver.from_nice_obj(pdd)
    A = dolfin.PETScMatrix()
    b = dolfin.PETScVector()
    solver.assembler.assemble(A)
    solver.assembler.assemble(b, solver.solution.u_.vector())
    indptr, indices, data = A.mat().getValuesCSR()
    W = solver.W
    mesh = W.mesh()
    cell_dofs = np.array([W.dofmap().cell_dofs(c.index()) for c in dolfin.cells(mesh)])
    np.savez_compressed(out, coordinates=mesh.coordinates(), cells=mesh.cells(),
                        cell_dofs=cell_dofs, indptr=indptr, indices=indices, data=data,
                        b=b.get_local(), u=solver.solution.u_.vector().get_local())


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'dolfin_reference.npz')

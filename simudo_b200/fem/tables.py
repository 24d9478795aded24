"""Pack the reference-element tables into the ``SbTables`` blob
(``simudo_b200/csrc/pdd_tables.h``) consumed by ``sb_plan_create``."""
import ctypes as C

import numpy as np

from .reference_element import reference_tables, collapsed_rule, dubiner_eval, \
    p2_lagrange_eval

QM = 11
QG = QM * QM


class SbRule(C.Structure):
    _fields_ = [('m', C.c_int32), ('pad_', C.c_int32),
                ('a', C.c_double * QM),
                ('wag', (C.c_double * 5) * QM),
                ('g3', (C.c_double * 3) * QM),
                ('b', C.c_double * QM),
                ('wbh', (C.c_double * 15) * QM),
                ('h6', (C.c_double * 6) * QM)]


class SbTables(C.Structure):
    _fields_ = [('bdmC', ((C.c_double * 6) * 2) * 12),
                ('G22u', (C.c_double * 15) * 21),
                ('G21', ((C.c_double * 10) * 3) * 6),
                ('G11', ((C.c_double * 6) * 3) * 3),
                ('G1', (C.c_double * 3) * 3),
                ('Mhat', ((C.c_double * 3) * 12) * 12),
                ('Dhat', (C.c_double * 3) * 12),
                ('M1', (C.c_double * 3) * 3),
                ('trI', C.c_double * 3),
                ('pad_', C.c_double),
                ('D3', ((C.c_double * 20) * 3) * 12),
                ('rho', SbRule), ('sup', SbRule),
                ('ng', C.c_int32), ('pad2_', C.c_int32),
                ('gX', C.c_double * QG), ('gY', C.c_double * QG),
                ('gwpsi', (C.c_double * 6) * QG),
                ('gp2', (C.c_double * 6) * QG)]


def _set(dst, arr):
    a = np.ascontiguousarray(arr, dtype=np.float64)
    view = np.ctypeslib.as_array(dst)
    view[tuple(slice(0, n) for n in a.shape)] = a


def _pack_rule(rule_struct, T, m):
    r = T.rule(m)
    rule_struct.m = m
    _set(rule_struct.a, r['a'])
    _set(rule_struct.wag, (r['wa'][None, :] * r['g'][:5]).T)
    _set(rule_struct.g3, r['g'][:3].T)
    _set(rule_struct.b, r['b'])
    _set(rule_struct.wbh, (r['wb'][None, :] * r['h']).T)
    _set(rule_struct.h6, r['h'][:6].T)


def pack_tables(model):
    """Return an ``SbTables`` instance for the quadrature sizes in ``model``."""
    T = reference_tables()
    t = SbTables()
    _set(t.bdmC, T.bdm_C)
    iu = [(r, s) for r in range(6) for s in range(r, 6)]
    _set(t.G22u, np.stack([T.G22[r, s] for r, s in iu]))
    _set(t.G21, T.G21)
    _set(t.G11, T.G11)
    _set(t.G1, T.G1)
    M = T.bdm_mass
    _set(t.Mhat, np.stack([M[:, :, 0, 0], M[:, :, 0, 1] + M[:, :, 1, 0], M[:, :, 1, 1]], axis=2))
    _set(t.Dhat, T.bdm_div_p1)
    _set(t.M1, T.p1_mass)
    _set(t.trI, [2.0 / 3.0, -1.0 / 3.0, 2.0 / 3.0])
    _set(t.D3, np.einsum('iar,rlt->ilat', T.bdm_C, T.G21).reshape(12, 3, 20))
    _pack_rule(t.rho, T, model.quad_m_rho)
    _pack_rule(t.sup, T, model.quad_m_super)
    pts, w = collapsed_rule(model.quad_m_g)
    t.ng = len(w)
    _set(t.gX, pts[:, 0])
    _set(t.gY, pts[:, 1])
    _set(t.gwpsi, (w[None, :] * dubiner_eval(2, pts)).T)
    _set(t.gp2, p2_lagrange_eval(pts).T)
    return t

"""Minimal ``dolfin`` stand-in for user scripts.

Simudo scripts ``import dolfin`` only to create the swept bias as
``U.V * dolfin.Constant(0.0)`` (``example/pn_diode/pn_diode.py:26,177``) and
the steppers mutate it with ``.assign`` (``fem/adaptive_stepper.py:331-334``).
``from simudo_b200 import dolfin`` keeps those lines working.
"""
from .fem.expr import Constant

__all__ = ['Constant']

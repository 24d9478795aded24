"""Developer helper: native-path parity under the CPU emulation (tests/emul)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import conftest
from simudo_b200 import _lib
so = conftest.build_emulation_library(); _lib._use_library_for_tests(so)
import hotpath_cases as hc
from simudo_b200 import synthetic as syn
cases = ((syn.ib_problem, 6, 3, True), (syn.pn_problem, 5, 3, True), (syn.ib_problem, 5, 4, False),
         (syn.pn_problem, 34, 3, True))
import sys
pats = sys.argv[1:] or ['native', 'bsr3']
for pat in pats:
    for mk, nx, ny, bcs in cases:
        try:
            pa, st, plan, vals, b, vals0, b0 = hc.assembly_parity(mk, nx, ny, numbering='b200', with_bcs=bcs, pattern=pat)
            print(pat, mk.__name__, nx, ny, bcs, 'native=', plan.is_native(), 'ok')
        except AssertionError as e:
            print(pat, mk.__name__, nx, ny, bcs, 'FAIL', e)

"""Generates the golden fixtures under tests/golden/ (run from the repo root:
``python tests/golden/make_golden.py``).

* ``tutorial_pn_diode_JV_svg.json``: the reference's own published tutorial
  J-V curve, decoded from ``/root/reference/doc/source/_static/
  tutorial-pn-diode-JV.svg`` (matplotlib path ``line2d_33``; decoding in
  SURVEY.md Appendix C: V = 0.2 + 0.8 (x - 153.419633) / 271.462293,
  log10 J = -4 + 8 (234.744452 - y) / 196.681814).  The only end-to-end number
  of the hot path anywhere in the reference tree.
* ``pn_diode_JV_oracle.json``: the same sweep computed with the oracle backend
  (C restatement + scipy LU), the fixture the GPU run is compared with.
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def decode_svg(path='/root/reference/doc/source/_static/tutorial-pn-diode-JV.svg'):
    txt = open(path).read()
    m = re.search(r'<g id="line2d_33">.*?<path[^>]* d="([^"]+)"', txt, flags=re.S)
    nums = re.findall(r'[ML]\s*([-\d.]+)\s+([-\d.]+)', m.group(1))
    out = []
    for xs, ys in nums:
        x, y = float(xs), float(ys)
        V = 0.2 + 0.8 * (x - 153.419633) / 271.462293
        J = 10 ** (-4 + 8 * (234.744452 - y) / 196.681814)
        out.append((round(V, 6), J))
    return out


def main():
    if os.path.exists('/root/reference'):
        jv = decode_svg()
        with open(os.path.join(HERE, 'tutorial_pn_diode_JV_svg.json'), 'w') as f:
            json.dump(dict(source='simudo doc/source/_static/tutorial-pn-diode-JV.svg line2d_33',
                           units=['V', 'mA/cm^2'], jv=jv), f, indent=1)
    from oracle_backend import OracleBackend
    from simudo_b200.example import pn_diode
    jv = pn_diode.run(backend_factory=OracleBackend)
    with open(os.path.join(HERE, 'pn_diode_JV_oracle.json'), 'w') as f:
        json.dump(dict(source='oracle backend, tests/golden/make_golden.py',
                       units=['V', 'mA/cm^2'], jv=[(float(v), float(j)) for v, j in jv]),
                  f, indent=1)


if __name__ == '__main__':
    main()

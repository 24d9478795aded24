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
* ``ib_cell_JV_oracle.json``: illuminated four-layer intermediate-band cell
  (example/fourlayer.py at the reduced size FOURLAYER_SMALL: light ramp +
  bias sweep with self-consistent optics) through the oracle backend.
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


# reduced fourlayer (BASELINE config 2 shape: 4 layers, 3 bands, 3 optical fields, X = 100)
FOURLAYER_SMALL = dict(p_thickness=0.2, n_thickness=0.2, IB_thickness=0.4, mesh_start=0.01,
                       mesh_factor=1.5, mesh_max=0.05, V_exts=[0.0, 0.2, 0.4], X=100,
                       CB_E=1.1, IB_E=0.65)


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
    from simudo_b200.example import fourlayer
    jv = fourlayer.run(P=FOURLAYER_SMALL, backend_factory=OracleBackend)
    with open(os.path.join(HERE, 'ib_cell_JV_oracle.json'), 'w') as f:
        json.dump(dict(source='oracle backend, tests/golden/make_golden.py, FOURLAYER_SMALL',
                       params={k: v for k, v in FOURLAYER_SMALL.items()},
                       units=['V', 'mA/cm^2'], jv=[(float(v), float(j)) for v, j in jv]),
                  f, indent=1)


def reference_parameter_fixture():
    """Config 2 at the reference's own parameters and mesh (fourlayer_example.py:56-110): the
    full light ramp + 26-point bias sweep through the oracle backend (7-8 minutes on 8 cores);
    the efficiency of this J-V is the number the tutorial quotes (34.1 %, tutorial.rst:60)."""
    import time
    import numpy as np
    from oracle_backend import OracleBackend
    from simudo_b200.example import fourlayer
    from simudo_b200.example.iv_analysis import IV_params, efficiency
    V = [float(v) for v in np.linspace(0, 1.7, 26)]
    t = time.time()
    jv = [(float(v), float(j)) for v, j in fourlayer.run(P=dict(V_exts=V), backend_factory=OracleBackend)]
    p = IV_params([j for _, j in jv], [v for v, _ in jv], -1.0)
    with open(os.path.join(HERE, 'ib_cell_reference_parameters_oracle.json'), 'w') as f:
        json.dump(dict(source='oracle backend, tests/golden/make_golden.py: fourlayer.run at the '
                              'defaults of the reference example, 26 voltages 0 .. 1.7 V; %.0f s'
                              % (time.time() - t),
                       units=['V', 'mA/cm^2'], jv=jv, efficiency=efficiency(jv), voc=p.voc,
                       jsc=p.isc, ff=p.ff, tutorial_efficiency=0.341,
                       tutorial_source='doc/source/user/tutorial.rst:60'), f, indent=1)


if __name__ == '__main__':
    if '--reference-parameters' in sys.argv:
        reference_parameter_fixture()
    else:
        main()

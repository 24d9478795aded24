import cProfile, pstats, io, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from simudo_b200.example import ibsc_ensemble as E, fourlayer
m = dict(E.members(1)[0], V_exts=list(np.linspace(0, 0.9, 6)))
import os
from simudo_b200.fem import linear_solver
if os.environ.get("MF"): linear_solver.BAND_LIMIT_BYTES = int(os.environ["MF"])
fourlayer.run(P=dict(m, V_exts=[0.0]))     # warm
pr = cProfile.Profile(); t = time.time(); pr.enable()
fourlayer.run(P=m)
pr.disable(); print('wall', time.time() - t)
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45); print(s.getvalue()[:9000])

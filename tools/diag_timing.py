import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpedm_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'build', 'libgpedm_timing.so')
import gpedm_b200 as G
from gpedm_b200 import synth
from gpedm_b200.batch import default_initparst
Y, X = synth.embed(synth.ricker(128 + 3, 1004), 3, 1)
m = G.GPModel(Y, X); p0 = default_initparst(3, True)
for _ in range(3): m.likegrad(p0, 1.0, None, None, True)
ts = (C.c_longlong * 32)()
G.lib().gpedm_debug_diag_timing(ts)
t = np.array(ts[:19]); print('cycles since start:', (t - t[0]).tolist()); print('deltas:', np.diff(t).tolist())
names = ['load','chol0','panel0','trail0','chol1','panel1','trail1','chol2','panel2','trail2','chol3','panel3','trail3','storeL+logdet','inv32','lvl1','lvl2','storeD']
print({n: int(d) for n, d in zip(names, np.diff(t))})

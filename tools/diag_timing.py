import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpedm_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'build', os.environ.get('GPEDM_TIMING_LIB', 'libgpedm_timing.so'))
import gpedm_b200 as G
from gpedm_b200 import synth
from gpedm_b200.batch import default_initparst
Y, X = synth.embed(synth.ricker(128 + 3, 1004), 3, 1)
m = G.GPModel(Y, X); p0 = default_initparst(3, True)
for _ in range(3): m.likegrad(p0, 1.0, None, None, True)
ts = (C.c_longlong * 32)()
G.lib().gpedm_debug_diag_timing(ts)
t = np.array(ts[:19], dtype=np.int64)
idx = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 14, 15, 17, 18]
names = ['load', 'chol0', 'panel0', 'diagupd0', 'chol1+trail0', 'panel1', 'diagupd1', 'chol2+trail1', 'panel2', 'diagupd2',
         'chol3', 'storeL+logdet', 'inv8', 'doubling', 'storeD']
tt = t[idx]
print({n: int(d) for n, d in zip(names, np.diff(tt))}, 'total', int(tt[-1] - tt[0]))

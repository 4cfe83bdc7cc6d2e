"""gpedm_b200: B200-native (sm_100a, FP64) GP fit / predict hot path of GPEDM behind the
reference's own API names.  All numerics run in libgpedm_b200.so (hand-written CUDA, C ABI in
include/gpedm.h); there is no CPU fallback."""
from ._lib import GPEDMError, build, lib  # noqa: F401
from .core import (GPModel, fmingrad_Rprop, fp64_peak, getcov, getcovinv, getlikegrad, logit)  # noqa: F401
from .fitgp import (GP, fitGP, getconditionals, getR2, getR2pop, getrhomatrix, makelags, predict,  # noqa: F401
                    predict_iter, predict_seq, summary, update)
from .batch import b_profile, fit_batch, fit_grid, multistart  # noqa: F401

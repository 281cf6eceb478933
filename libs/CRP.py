"""Drop-in for the reference module libs/CRP.py: same class name, constructor and
method contract (cbg-ethz/BnpC libs/CRP.py:17-66; callers run_BnpC.py:251-255,
libs/MCMC.py:128-135,242-282,320-342), backed by the sm_100a kernels of
bnpc_b200.  Nothing here computes on the CPU."""
import numpy as np

from bnpc_b200.engine import EPS, DeviceCRP

EPSILON = EPS
TMIN = 1e-5
TMAX = 1 - TMIN
log_EPSILON = np.log(EPSILON)


class CRP(DeviceCRP):
    """
    Arguments:
        data (np.array): n x m matrix with n cells and m mutations containing 0|1|np.nan
        DP_alpha ((float, float)): Gamma prior of the CRP concentration (either < 0: (sqrt(n), 1))
        param_beta ((float, float)): Beta prior of the cluster parameters
        FN_error (float): fixed false negative rate
        FP_error (float): fixed false positive rate
    Extra keyword arguments: device (torch device), rnd (bnpc_b200.rng source).
    """

    def __init__(self, data, DP_alpha=(-1, -1), param_beta=(1, 1), FN_error=EPSILON,
                 FP_error=EPSILON, **kw):
        super().__init__(data, DP_alpha, param_beta, FN_error, FP_error, **kw)

"""Drop-in for the reference module libs/CRP_learning_errors.py (class
`CRP_errors_learning`, libs/CRP_learning_errors.py:17-32).  The chain driver
recognises the learning model by this module path (libs/MCMC.py:206), so the path
and class name are kept."""
from bnpc_b200.engine import DeviceCRPLearnErrors


class CRP_errors_learning(DeviceCRPLearnErrors):
    def __init__(self, data, DP_alpha=(-1, -1), param_beta=(1, 1), FP_mean=0.001, FP_sd=0.0005,
                 FN_mean=0.25, FN_sd=0.05, **kw):
        super().__init__(data, DP_alpha, param_beta, FP_mean, FP_sd, FN_mean, FN_sd, **kw)

"""bnpc_b200 -- B200-native (sm_100a) implementation of the BnpC per-step MCMC hot path.

Public surface:
  bnpc_b200.engine.DeviceCRP / DeviceCRPLearnErrors   CUDA-backed models (reference class contract)
  bnpc_b200.rng.PhiloxRandom / TapeRandom             random sources (production / parity replay)
  bnpc_b200._lib.build()                              compile libbnpc_b200.so in-tree with nvcc

The drop-in module paths of the reference (`libs.CRP.CRP`,
`libs.CRP_learning_errors.CRP_errors_learning`, `libs.MCMC.MCMC`) live in the top-level
`libs/` package and delegate here.
"""
import os as _os

# chains run concurrently on one stream each; with the default 8 hardware queues short kernels
# of one chain would queue behind another chain's long sweep (must be set before CUDA starts)
_os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')

__version__ = '0.1.0'

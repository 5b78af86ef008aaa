"""apemost_b200 -- B200-native parallel-tempering MCMC engine for APEMoST's hot path.

The product is ``libapemost_gpu.so`` (CUDA sm_100a behind the C ABI in
``include/apemost_gpu.h``) plus the C host layer in ``apemost_b200/host``.
``apemost_b200.capi`` is the ctypes binding the tests and bench.py use.
"""
__version__ = "0.1.0"

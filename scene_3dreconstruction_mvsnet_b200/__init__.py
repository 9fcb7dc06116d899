"""B200-native (sm_100a) MVSNet depth-inference hot path.

Package layout: csrc/ (CUDA kernels + C ABI), _lib.py (ctypes binding), ops.py (torch tensors ->
device pointers, autograd plumbing), models/ (host-side mirror of the reference's models package),
sharding.py (reference-view partitioning across ranks), synth.py (seeded synthetic inputs).
"""
__version__ = "0.1.0"

"""
quantum_inferno_b200 -- B200 (sm_100a) implementation of the quantum-inferno FFT time-frequency hot path.

Drop-in module names for the reference path (``styx_cwt``, ``styx_stx``, ``styx_fft``, ``cwt_atoms``,
``tfr_info``, ``scales_dyadic``); the arithmetic runs in hand-written CUDA kernels behind the C ABI of
``libqi_b200.so`` (include/qi_b200.h).  There is no CPU fallback: importing a transform module works
anywhere, calling it without the compiled library or without a CUDA device raises.
"""
__version__ = "0.1.0"

"""
Device plumbing: PyTorch is used only to own HBM buffers and the CUDA stream; every kernel is launched
through the C ABI of libqi_b200.so with raw pointers.

``get_runtime()`` returns the CUDA runtime or raises -- there is no CPU path in this package.
(The test-suite injects its own debug runtime with ``use_runtime`` to drive a g++ build of the same
kernel sources on machines without a GPU; that object lives under tests/emul/, not here.)
"""
import contextlib

import numpy as np

from . import _lib

DTYPE_CODE = {"float32": _lib.QI_F32, "float64": _lib.QI_F64}
COMPLEX_OF = {"float32": "complex64", "float64": "complex128"}


def dtype_name(dtype, default="float64"):
    """Normalise a user dtype (None, str, numpy or torch dtype) to 'float32' / 'float64'."""
    if dtype is None:
        return default
    s = str(dtype).replace("torch.", "")
    s = {"f4": "float32", "f8": "float64", "single": "float32", "double": "float64", "fp32": "float32",
         "fp64": "float64", "<class 'numpy.float32'>": "float32", "<class 'numpy.float64'>": "float64"}.get(s, s)
    if s not in DTYPE_CODE:
        raise ValueError(f"dtype must be float32 or float64, got {dtype!r}")
    return s


class _DeviceBoundLib:
    """The bound C library with every ``qi_*`` call made while ``device`` is the current CUDA device: the kernels are
    launched with ``<<<>>>`` on the stream handed in, so the device that owns that stream and the buffers has to be
    current whatever the caller's ``torch.cuda.current_device()`` is."""

    def __init__(self, lib, torch, device):
        self._lib, self._torch, self._device = lib, torch, device
        self._cache = {}

    def __getattr__(self, name):
        fn = self._cache.get(name)
        if fn is None:
            raw, guard, dev = getattr(self._lib, name), self._torch.cuda.device, self._device

            def fn(*args):
                with guard(dev):
                    return raw(*args)

            self._cache[name] = fn
        return fn


class CudaRuntime:
    """torch-backed buffers on ONE CUDA device + the bound CUDA library.

    One runtime exists per device (``get_runtime(like=tensor)`` picks the tensor's device).  Scratch memory is one
    grow-only buffer per (device, stream): calls on the same stream are ordered, so they can share it; calls on
    different streams get different buffers.  Host threads must not share a stream (plain CUDA stream semantics)."""
    name = "cuda"

    def __init__(self, device=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("quantum_inferno_b200 needs a CUDA device (B200, sm_100a); none is visible and "
                               "there is no CPU fallback")
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib = _DeviceBoundLib(_lib.load(), torch, self.device)
        self._ws = {}

    # ---- buffers
    def _tdtype(self, name):
        return getattr(self.torch, name)

    def empty(self, shape, dtype):
        return self.torch.empty(tuple(int(s) for s in shape), dtype=self._tdtype(dtype), device=self.device)

    def zeros(self, shape, dtype):
        return self.torch.zeros(tuple(int(s) for s in shape), dtype=self._tdtype(dtype), device=self.device)

    def is_device_array(self, x):
        return isinstance(x, self.torch.Tensor)

    def asarray(self, x, dtype):
        """numpy / torch (any device) -> contiguous tensor of `dtype` on this device."""
        if not isinstance(x, self.torch.Tensor):
            x = np.ascontiguousarray(x)
            x = self.torch.from_numpy(x if x.flags.writeable else x.copy())    # read-only arrays (npz, scipy windows)
        t = x
        return t.to(device=self.device, dtype=self._tdtype(dtype), non_blocking=True).contiguous()

    def ptr(self, buf):
        return 0 if buf is None else buf.data_ptr()

    def stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def workspace(self, nbytes):
        """Grow-only scratch buffer of the current stream (256-byte aligned by the caching allocator).  A buffer that
        is outgrown goes back to the caching allocator, which keeps it alive until the stream has passed its last use."""
        key = self.stream()
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            self._ws.pop(key, None)
            with self.torch.cuda.device(self.device):
                ws = self.torch.empty(int(nbytes), dtype=self.torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    def release_workspace(self):
        """Give every scratch buffer back to the caching allocator."""
        self._ws.clear()

    def to_numpy(self, buf):
        return buf.detach().cpu().numpy()

    def randn(self, shape, dtype, seed=None):
        """Standard normal samples from the device generator; ``seed=None`` draws fresh entropy."""
        gen = self.torch.Generator(device=self.device)
        if seed is None:
            gen.seed()
        else:
            gen.manual_seed(int(seed))
        return self.torch.randn(tuple(int(s) for s in shape), dtype=self._tdtype(dtype), device=self.device, generator=gen)

    def reshape(self, buf, shape):
        return buf.reshape(tuple(shape))


_runtime = None          # override installed by use_runtime() (test-suite emulator); wins over the per-device table
_cuda_runtimes = {}      # device index -> CudaRuntime


def get_runtime(like=None):
    """The runtime of the device that holds ``like`` (a CUDA tensor), else of torch's current CUDA device."""
    if _runtime is not None:
        return _runtime
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("quantum_inferno_b200 needs a CUDA device (B200, sm_100a); none is visible and there is "
                           "no CPU fallback")
    if isinstance(like, torch.Tensor) and like.is_cuda:
        index = like.device.index
    else:
        index = torch.cuda.current_device()
    rt = _cuda_runtimes.get(index)
    if rt is None:
        rt = _cuda_runtimes[index] = CudaRuntime(torch.device("cuda", index))
    return rt


@contextlib.contextmanager
def use_runtime(rt):
    """Temporarily install another runtime object (used by the test-suite's kernel-logic emulator)."""
    global _runtime
    prev, _runtime = _runtime, rt
    try:
        yield rt
    finally:
        _runtime = prev


def finish(rt, buf, want_numpy):
    return rt.to_numpy(buf) if want_numpy else buf

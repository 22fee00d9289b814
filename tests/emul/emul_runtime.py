"""TEST INFRASTRUCTURE: a numpy-backed stand-in for quantum_inferno_b200._runtime.CudaRuntime that drives
tests/emul/libqi_emul.so (the g++ -DQI_EMUL build of the kernel sources).  It exists so the host-side
drivers and the kernel logic can be exercised on a machine without a GPU; it is injected with
``_runtime.use_runtime`` by the tests and is not reachable from the package itself."""
import numpy as np

from . import emul_lib


class EmulRuntime:
    name = "emul"

    def __init__(self):
        self.lib = emul_lib.load()

    @staticmethod
    def _aligned(shape, dtype, align=256):
        """Like a device allocation: the first byte sits on a 256-byte boundary (the C ABI checks plane alignment)."""
        shape = tuple(int(s) for s in shape)
        dt = np.dtype(dtype)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dt.itemsize
        raw = np.empty(nbytes + align, dtype=np.uint8)
        off = (-raw.ctypes.data) % align
        return raw[off:off + nbytes].view(dt).reshape(shape)

    def empty(self, shape, dtype):
        return self._aligned(shape, dtype)

    def zeros(self, shape, dtype):
        out = self._aligned(shape, dtype)
        out[...] = 0
        return out

    def is_device_array(self, x):
        return False

    def asarray(self, x, dtype):
        out = self._aligned(np.shape(x), dtype)
        out[...] = np.asarray(x)
        return out

    def ptr(self, buf):
        return None if buf is None else buf.ctypes.data

    def stream(self):
        return None

    def workspace(self, nbytes):
        raw = np.empty(int(nbytes) + 256, dtype=np.uint8)
        off = (-raw.ctypes.data) % 256
        self._ws_keep = raw
        return raw[off:off + int(nbytes)]

    def to_numpy(self, buf):
        return np.asarray(buf)

    def randn(self, shape, dtype, seed=None):
        return np.random.default_rng(seed).standard_normal(tuple(int(s) for s in shape)).astype(dtype)

    def reshape(self, buf, shape):
        return buf.reshape(tuple(shape))

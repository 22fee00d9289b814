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

    def empty(self, shape, dtype):
        return np.empty(tuple(int(s) for s in shape), dtype=dtype)

    def zeros(self, shape, dtype):
        return np.zeros(tuple(int(s) for s in shape), dtype=dtype)

    def is_device_array(self, x):
        return False

    def asarray(self, x, dtype):
        return np.ascontiguousarray(np.asarray(x), dtype=dtype)

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

    def reshape(self, buf, shape):
        return buf.reshape(tuple(shape))

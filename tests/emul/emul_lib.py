"""TEST INFRASTRUCTURE: builds/loads tests/emul/libqi_emul.so (the g++ -DQI_EMUL build of the kernel sources)."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def load():
    global _lib
    if _lib is None:
        from quantum_inferno_b200._lib import bind
        override = os.environ.get("QI_EMUL_LIB")                # e.g. the AddressSanitizer build of asan_check.sh
        if override:
            _lib = bind(ctypes.CDLL(override))
            return _lib
        subprocess.run([os.path.join(HERE, "build_emul.sh")], check=True, capture_output=True)
        _lib = bind(ctypes.CDLL(os.path.join(HERE, "libqi_emul.so")))
    return _lib

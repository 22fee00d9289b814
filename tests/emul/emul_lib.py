"""TEST INFRASTRUCTURE: builds/loads tests/emul/libqi_emul.so (the g++ -DQI_EMUL build of the kernel sources)."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def load():
    global _lib
    if _lib is None:
        subprocess.run([os.path.join(HERE, "build_emul.sh")], check=True, capture_output=True)
        from quantum_inferno_b200._lib import bind
        _lib = bind(ctypes.CDLL(os.path.join(HERE, "libqi_emul.so")))
    return _lib

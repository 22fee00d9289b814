"""
CPU-emulator evaluation of a level-assignment setting of the multirate CWT (`QI_MR_KAPPA`, `QI_MR_ENV_KAPPA`) against the
assertions of the `-m gpu` tier: the kernel sources run through tests/emul (test infrastructure; same arithmetic as the
device up to FMA contraction and lg2.approx), the numpy oracle is the checker.

    python tools/level_assignment_eval.py                                   # the defaults
    QI_MR_KAPPA=1.5 QI_MR_ENV_KAPPA=2.4 python tools/level_assignment_eval.py [24]   # optional: one 2^24 record too

Prints, per configuration of tests/test_gpu_parity.py::test_multirate_vs_oracle / _properties_ / _other_orders, the
figures those tests bound (plane and per-band L2, entropy, total power, pdf sum, information plane).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from quantum_inferno_b200 import _plan, _runtime, cwt_entropy  # noqa: E402
from tests.emul.emul_runtime import EmulRuntime  # noqa: E402
from tests.test_gpu_parity import l2, synth  # noqa: E402
from oracle import qi_oracle as orc  # noqa: E402

FS = 800.0
_ctx = _runtime.use_runtime(EmulRuntime())
_ctx.__enter__()
print("MR_KAPPA", _plan.MR_KAPPA, "QI_MR_ENV_KAPPA", os.environ.get("QI_MR_ENV_KAPPA", "4.8 (default)"), flush=True)

for order, logn in ((3, 13), (3, 16), (6, 14), (12, 13), (1, 13), (3, 18)):
    n = 1 << logn
    x = np.stack([synth(n, chan=0), synth(n, chan=5)[::-1].copy()])[:1 if logn >= 18 else 2]
    r = cwt_entropy.cwt_power_entropy(order, x, FS, dtype="float32", method="multirate")
    for c in range(x.shape[0]):
        ref = orc.cwt_power_entropy(order, x[c], FS)
        p = np.asarray(r.power[c], dtype=np.float64)
        per = np.linalg.norm(p - ref["power"], axis=1) / np.linalg.norm(ref["power"], axis=1)
        strong = ref["power"] > 1e-2 * ref["power"].max()
        print(f"vs oracle  order {order} 2^{logn} ch {c}: L2 {l2(p, ref['power']):.2e} (2e-5)  per-band {per.max():.2e} (5e-5)  "
              f"entropy {abs(float(r.entropy_bits()[c]) - ref['entropy_bits']):.1e} (1e-4)  "
              f"total {abs(float(r.total_power[c]) - ref['total']) / ref['total']:.1e} (1e-5)  "
              f"pdf sum - 1 {(p / float(r.total_power[c])).sum() - 1:.1e} (2e-6)  "
              f"info {np.abs(np.asarray(r.info[c], dtype=np.float64) - ref['info'])[strong].max():.1e} (1e-3)", flush=True)

n, C = 1 << 20, 4
x = np.stack([synth(n, chan=c) for c in range(C)])
r = cwt_entropy.cwt_power_entropy(3, x, FS, dtype="float32")
p = np.asarray(r.power, dtype=np.float64)
print("properties 4 x 2^20: pdf sum - 1 per channel", (p / np.asarray(r.total_power)[:, None, None]).sum((1, 2)) - 1, "(2e-6)", flush=True)

for order, logn in ((6, 20), (12, 18), (1.5, 19)):
    n = 1 << logn
    x = synth(n, chan=1)[None, :]
    a = cwt_entropy.cwt_power_entropy(order, x, FS, dtype="float32", method="multirate")
    b = cwt_entropy.cwt_power_entropy(order, x, FS, dtype="float64", method="exact", want_info=True)
    pa, pb = np.asarray(a.power[0], dtype=np.float64), np.asarray(b.power[0], dtype=np.float64)
    per = np.linalg.norm(pa - pb, axis=-1) / np.linalg.norm(pb, axis=-1)
    print(f"vs exact   order {order} 2^{logn}: per-band {per.max():.2e} (1e-4)  L2 {np.linalg.norm(pa - pb) / np.linalg.norm(pb):.2e} (2e-5)  "
          f"entropy {abs(float(a.entropy_bits()[0]) - float(b.entropy_bits()[0])):.1e} (1e-4)  "
          f"total {abs(float(a.total_power[0]) - float(b.total_power[0])) / float(b.total_power[0]):.1e} (1e-5)", flush=True)

if len(sys.argv) > 1:
    logn = int(sys.argv[1])
    n = 1 << logn
    x = synth(n, chan=3)[None, :]
    t0 = time.time()
    r = cwt_entropy.cwt_power_entropy(3, x, FS, dtype="float32", method="multirate")
    p = np.asarray(r.power[0], dtype=np.float64)
    print(f"north-star record 2^{logn} ({time.time() - t0:.0f} s): pdf sum - 1 {(p / float(r.total_power[0])).sum() - 1:.2e} (2e-6)", flush=True)
    xf = np.fft.fft(x[0].astype(np.float64), 2 * n)
    nb = p.shape[0]
    for band in (0, 1, 10, 35 if nb > 35 else nb // 2, nb - 1):
        row = np.abs(orc.cwt_band(xf, 3, n, r.frequency_hz[band], FS)) ** 2
        print(f"  band {band}: L2 {l2(p[band], row):.2e} (1e-4)", flush=True)

"""
Host-side filter bookkeeping for csrc/qi_iir.cu (a few dozen numbers per call).

The reference filters in transfer-function form: ``scipy.signal.filtfilt(b, a, x)`` with the (b, a) that
``scipy.signal.butter`` rounded to float64 (styx_fft.py:86-90, synth/synthetic_signals.py:190-192).  A blocked parallel
evaluation of that direct-form recursion would lose accuracy -- its state transition matrix is highly non-normal
(transient growth G ~ 1e4..1e7 for the reference's band-passes), and splitting the recursion into blocks turns the
sequential eps*G rounding noise into eps*G^2.  The SAME transfer function is therefore evaluated as a cascade of
second-order sections: the zeros and poles of the ROUNDED polynomials (not of the ideal Butterworth design) are found
with 120-digit arithmetic (mpmath), paired into real sections, and the steady-state initial conditions follow
``scipy.signal.sosfilt_zi``.  In exact arithmetic the cascade started from its steady state equals
``lfilter(b, a, x, zi=lfilter_zi(b, a) * x[0])``; numerically it is closer to that exact value than the
direct form is (tests/test_iir_*: long-double restatement of the reference recursion as the arbiter).
"""
import numpy as np

MAX_STATE = 16


def _real_sections(roots, tol):
    """Group the roots of a real polynomial into (r1, r2) pairs: conjugates together, reals with reals; a last lone
    real root is returned as (r, None)."""
    cplx = sorted([r for r in roots if abs(r.imag) > tol and r.imag > 0], key=lambda r: (r.real, r.imag))
    real = sorted([r.real for r in roots if abs(r.imag) <= tol])
    n_neg = sum(1 for r in roots if abs(r.imag) > tol and r.imag < 0)
    if n_neg != len(cplx):
        raise ValueError("roots of a real polynomial must come in conjugate pairs")
    pairs = [(r, r.conjugate()) for r in cplx]
    while len(real) >= 2:
        pairs.append((real.pop(0), real.pop(0)))
    if real:
        pairs.append((real[0], None))
    return pairs


_SOS_CACHE = {}


def tf2sos_exact(b, a):
    """Cached ``_tf2sos_exact`` (the 120-digit factorisation costs ~0.1-1 s of host time per new filter)."""
    key = (np.asarray(b, dtype=np.float64).tobytes(), np.asarray(a, dtype=np.float64).tobytes())
    if key not in _SOS_CACHE:
        if len(_SOS_CACHE) > 256:
            _SOS_CACHE.clear()
        _SOS_CACHE[key] = _tf2sos_exact(b, a)
    return _SOS_CACHE[key].copy()


def _tf2sos_exact(b, a):
    """(b, a) float64 taps of equal length -> sos [n_sections, 6] float64 of the same transfer function.

    Unlike ``scipy.signal.tf2sos`` the roots are those of the given rounded polynomials to full double precision
    (120-digit companion eigenvalues), so the cascade realises the reference's own filter, not a neighbour of it."""
    import mpmath as mp
    b = np.atleast_1d(np.asarray(b, dtype=np.float64))
    a = np.atleast_1d(np.asarray(a, dtype=np.float64))
    n = max(len(a), len(b))
    b = np.concatenate((b, np.zeros(n - len(b)))) / a[0]
    a = np.concatenate((a, np.zeros(n - len(a)))) / a[0]
    if b[0] == 0.0 or n < 2 or n - 1 > MAX_STATE:
        raise ValueError("filter needs b[0] != 0 and 1 <= order <= %d" % MAX_STATE)
    with mp.workdps(120):
        tol = mp.mpf(10) ** -25

        def roots_of(c):
            c = [mp.mpf(float(v)) for v in c]
            n_zero = 0
            while len(c) > 1 and c[-1] == 0:             # trailing zero taps: roots at the origin
                c.pop()
                n_zero += 1
            r = []
            if len(c) == 2:
                r = [-c[1] / c[0]]
            elif len(c) > 2:                              # eigenvalues of the companion matrix (QR iteration copes
                deg = len(c) - 1                          # with the clustered zeros of (z - 1)^N, Durand-Kerner does not)
                comp = mp.zeros(deg, deg)
                for k in range(deg):
                    comp[0, k] = -c[k + 1] / c[0]
                    if k + 1 < deg:
                        comp[k + 1, k] = mp.mpf(1)
                r = mp.eig(comp, left=False, right=False)
            return [mp.mpc(v) for v in r] + [mp.mpc(0)] * n_zero

        zeros, poles = _real_sections(roots_of(b), tol), _real_sections(roots_of(a), tol)
        if len(zeros) != len(poles):
            raise ValueError("numerator and denominator must factor into the same number of sections")
        # sections are ordered by pole radius (least damped last, as scipy.signal.zpk2sos) and every pole pair takes the
        # zero pair nearest to it; a lone real pole takes the lone real zero
        poles.sort(key=lambda p: (p[1] is None, abs(p[0])))
        sos = np.zeros((len(poles), 6))
        for i, (p1, p2) in enumerate(poles):
            cand = [z for z in zeros if (z[1] is None) == (p2 is None)] or zeros
            z = min(cand, key=lambda zz: abs(zz[0] - p1))
            zeros.remove(z)
            z1, z2 = z
            sec_b = [mp.mpf(1), -(z1 + z2), z1 * z2] if z2 is not None else [mp.mpf(1), -z1, mp.mpf(0)]
            sec_a = [mp.mpf(1), -(p1 + p2), p1 * p2] if p2 is not None else [mp.mpf(1), -p1, mp.mpf(0)]
            sos[i, :3] = [float(mp.re(v)) for v in sec_b]
            sos[i, 3:] = [float(mp.re(v)) for v in sec_a]
        sos[0, :3] *= b[0]
    return sos


def sosfilt_zi(sos):
    """scipy.signal.sosfilt_zi: steady state of every section of the cascade for a unit step at its input."""
    from scipy import signal
    return signal.sosfilt_zi(np.asarray(sos, dtype=np.float64))

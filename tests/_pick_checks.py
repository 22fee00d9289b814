"""Shared assertions for the sampling / picker drop-ins: the same checks run through the CPU
emulator of the kernels (tests/test_emul_api.py) and on the B200 (tests/test_gpu_parity.py, -m gpu)."""
import numpy as np
import pytest

METHODS = ("average", "median", "max", "min", "nth")


def same(a, b):
    """Bit-exact, NaNs in the same places."""
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)


def check_subsample_golden(golden):
    from quantum_inferno_b200.utilities import sampling
    g = golden("pick")
    for dt, tol in (("float64", 1e-14), ("float32", 2e-6)):
        plane = g["plane"].astype(dt)
        for f in g["factors"]:
            for m in METHODS:
                want = g[f"sub2d_{dt}_{int(f)}_{m}"]
                got = sampling.subsample_2d(plane, int(f), m)
                if m == "average":       # fp64 accumulation here, numpy's pairwise sum in the array dtype there
                    assert got.shape == want.shape and got.dtype == want.dtype, (dt, f, m)
                    assert np.array_equal(np.isnan(got), np.isnan(want))
                    ok = ~np.isnan(want)
                    assert np.max(np.abs(got[ok] - want[ok]) / np.abs(want[ok])) < tol, (dt, f, m)
                else:
                    assert same(got, want), (dt, f, m)
    for f in (2, 5, 64, 200):
        for m in METHODS:
            y, rate = sampling.subsample(g["row"], 800.0, f, m)
            assert rate == float(g[f"sub1d_{f}_{m}_rate"])
            if m == "average":
                assert np.allclose(y, g[f"sub1d_{f}_{m}"], rtol=1e-13, atol=1e-15)
            else:
                assert same(y, g[f"sub1d_{f}_{m}"]), (f, m)


def check_subsample_edges(capsys):
    from quantum_inferno_b200.utilities import sampling
    x = np.arange(10.0)
    y, r = sampling.subsample(x, 8.0, 1, "max")                       # factor < 2: the input itself + warning
    assert y is x and r == 8.0 and "less than 2" in capsys.readouterr().out
    y, r = sampling.subsample(x, 8.0, 3, "bogus")                     # unknown method -> "nth" + warning
    assert same(y, x[::3]) and "not recognized" in capsys.readouterr().out
    assert sampling.subsample_2d(x[None, :], 0) is not None
    y, _ = sampling.subsample(x, 8.0, 20, "average")                  # factor longer than the record: empty
    assert y.shape == (0,)
    y, _ = sampling.subsample(x, 8.0, 20, "nth")
    assert same(y, x[:1])
    p3 = np.arange(2 * 3 * 12, dtype=np.float64).reshape(2, 3, 12)    # batch axis extension
    assert same(sampling.subsample_2d(p3, 4, "max"), p3.reshape(2, 3, 3, 4).max(axis=3))
    ints = np.arange(12).reshape(2, 6)                                 # integer input computes in float64
    assert np.array_equal(sampling.subsample_2d(ints, 2, "median"), np.median(ints.reshape(2, 3, 2), axis=2))
    with pytest.raises(ValueError):
        sampling.subsample(np.zeros((2, 2)), 1.0, 2)


def check_subsample_vs_oracle(rng_seed=3):
    """Long groups (warp-per-group path, radix-select median) and ragged tails against the numpy oracle."""
    from oracle import qi_oracle as orc
    from quantum_inferno_b200.utilities import sampling
    rng = np.random.default_rng(rng_seed)
    for dt in ("float32", "float64"):
        p = (rng.standard_normal((3, 10007)) * 10.0 ** rng.integers(-3, 4, size=(3, 10007))).astype(dt)
        p[0, 5000:5600] = 1.25                                          # heavy ties
        p[2, 777] = np.nan
        p[1, 100] = -np.inf
        p[1, 9000] = np.inf
        for f in (129, 256, 1000, 3333, 5003, 10007):
            for m in ("median", "max", "min", "nth"):
                assert same(sampling.subsample_2d(p, f, m), orc.subsample_2d(p, f, m)), (dt, f, m)
            a, b = sampling.subsample_2d(p, f, "average"), orc.subsample_2d(p, f, "average")
            ok = np.isfinite(b)
            assert np.array_equal(np.isfinite(a), ok)
            assert np.allclose(a[ok], b[ok], rtol=1e-5 if dt == "float32" else 1e-12)


def check_subsample_vector_path():
    """Power-of-two factors on 16-byte aligned rows take the register / shuffle kernel (subsample_vec_kernel)."""
    from oracle import qi_oracle as orc
    from quantum_inferno_b200.utilities import sampling
    rng = np.random.default_rng(9)
    for dt in ("float32", "float64"):
        p = rng.standard_normal((3, 5 * 1024 + 72)).astype(dt)
        p[0, 1000] = np.nan
        p[1, 2049] = np.inf
        p[2, 4097] = -np.inf
        p[2, 64:128] = -3.5
        for f in (2, 4, 8, 16, 32, 64, 128):
            for m in ("max", "min", "nth"):
                assert same(sampling.subsample_2d(p, f, m), orc.subsample_2d(p, f, m)), (dt, f, m)
            a, b = sampling.subsample_2d(p, f, "average"), orc.subsample_2d(p, f, "average")
            assert a.shape == b.shape and a.dtype == b.dtype
            assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.isinf(a), np.isinf(b))
            ok = np.isfinite(b)
            assert np.allclose(a[ok], b[ok], rtol=2e-6 if dt == "float32" else 1e-13, atol=1e-7 if dt == "float32" else 1e-15)


def check_picker_golden(golden):
    from quantum_inferno_b200.utilities import picker
    g = golden("pick")
    x = g["x"]
    for et in picker.EXTRACTION_TYPE:
        got = picker.scale_signal_by_extraction_type(x, et)
        assert np.max(np.abs(got - g[f"scaled_{et}"])) <= 4e-16 * np.max(np.abs(g[f"scaled_{et}"])), et
        for h in (0.7, 0.3):
            pk = picker.find_peaks_by_extraction_type(x, et, h)
            assert pk.dtype == np.int64 and np.array_equal(pk, g[f"peaks_{et}_{h}"]), (et, h)
    for st in picker.INPUT_SCALE_TYPE:
        for tb in (1, 3):
            for dist in (0.1, 0.01, 0.5):
                pk = picker.find_peaks_with_bits(x, 800.0, st, tb, dist)
                assert np.array_equal(pk, g[f"bits_{st}_{tb}_{dist}"]), (st, tb, dist)
    assert np.array_equal(picker.find_peaks_with_bits(g["noise"], 800.0, "log2", 2, 0.05), g["noise_peaks_bits"])
    assert np.array_equal(picker.find_peaks_by_extraction_type(g["noise"], "sigmax", 0.5), g["noise_peaks_sigmax"])
    # float32 record: same positions (the record is well conditioned)
    assert np.array_equal(picker.find_peaks_by_extraction_type(x.astype(np.float32), "sigmax", 0.7), g["peaks_sigmax_0.7"])


def check_picker_edges(capsys):
    from oracle import qi_oracle as orc
    from quantum_inferno_b200.utilities import picker
    flat = np.array([0.0, 1.0, 1.0, 1.0, 0.0, 2.0, 2.0, 0.0, 3.0, 3.0])   # plateaus; the last one touches the edge
    assert np.array_equal(picker.find_peaks_by_extraction_type(flat, "sigmax", 0.0), [2, 5])
    assert np.array_equal(orc.find_peaks(flat / 3.0, height=0.0), [2, 5])
    assert picker.find_peaks_by_extraction_type(np.ones(50), "sigmax", 0.5).size == 0
    assert picker.find_peaks_by_extraction_type(np.array([1.0, 2.0]), "sigmax", 0.5).size == 0
    nanrec = np.array([0.0, 1.0, 0.0, np.nan, 0.0, 5.0, 0.0, 1.0])
    assert np.array_equal(picker.find_peaks_by_extraction_type(nanrec, "sigmax", 0.5), [5])     # nanmax ignores NaN
    assert picker.find_peaks_with_bits(nanrec, 8.0, "amplitude", 1, 0.2).size == 0              # np.max -> NaN height
    picker.scale_signal_by_extraction_type(flat, "nope")
    assert "Invalid extraction type" in capsys.readouterr().out
    with pytest.raises(TypeError):
        picker.find_peaks_with_bits(flat, 8.0, "amplitude", 1, 0.2, 5)
    with pytest.raises(ValueError):
        picker.find_peaks_with_bits(flat, 8.0, "amplitude", 1, 0.0)        # distance 0: scipy's ValueError
    with pytest.raises(ValueError):
        picker.find_peaks_by_extraction_type(np.zeros((3, 3)))
    with pytest.raises(ValueError):
        picker.extract_signal_index_with_buffer(8.0, 4, -1.0, 1.0)
    assert picker.extract_signal_index_with_buffer(8.0, 40, 1.0, 2.0) == (32, 56)
    assert np.array_equal(picker.extract_signal_with_buffer_seconds(flat, 1.0, 5, 2.0, 2.0), flat[3:7])
    assert np.array_equal(picker.find_peaks_to_comb_function(flat, np.array([2, 5])), np.eye(10)[[2, 5]].sum(axis=0))


def check_picker_vs_oracle(n=200000):
    """Dense candidate lists (capacity regrowth), distance selection and plateaus on a quantised record."""
    from oracle import qi_oracle as orc
    from quantum_inferno_b200.utilities import picker
    rng = np.random.default_rng(11)
    x = np.round(rng.standard_normal(n) * 8.0) / 8.0 + 4.0 * np.sin(np.arange(n) * 2e-3)          # many exact ties
    for st, tb, dist in (("log2", 6, 0.001), ("amplitude", 2, 0.05), ("log2", 1, 1.0)):
        assert np.array_equal(picker.find_peaks_with_bits(x, 1000.0, st, tb, dist),
                              orc.find_peaks_with_bits(x, 1000.0, st, tb, dist)), (st, tb, dist)
    for et in ("sigmax", "sigabs", "log2max"):
        assert np.array_equal(picker.find_peaks_by_extraction_type(x, et, 0.2),
                              orc.find_peaks_by_extraction_type(x, et, 0.2)), et

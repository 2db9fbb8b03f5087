"""The fixed-point scatter-add (csrc/dfx_phases.h: fx_scatter / fx_value / fx_pow2_scale) in isolation, on the CPU,
against exact integer arithmetic: the two 32-bit words hold hi * 2^21 + lo == sum of rint(x * scale), whatever the
order of the contributions; values outside the representable range raise the poison bit instead of wrapping."""
import ctypes

import numpy as np
import pytest

from emu_util import build_emu

_F = ctypes.POINTER(ctypes.c_float)
_I = ctypes.POINTER(ctypes.c_int)


def _lib(es=1):
    lib = build_emu(es)
    lib.emu_fx_accumulate.argtypes = [_F, _I, ctypes.c_int, ctypes.c_float, _I, _I, _F]
    lib.emu_fx_pow2_scale.restype = ctypes.c_float
    lib.emu_fx_pow2_scale.argtypes = [ctypes.c_float]
    return lib


def _acc(lib, vals, perm, scale):
    vals = np.ascontiguousarray(vals, np.float32)
    perm = np.ascontiguousarray(perm, np.int32)
    lo, hi, value = ctypes.c_int(), ctypes.c_int(), ctypes.c_float()
    bad = lib.emu_fx_accumulate(vals.ctypes.data_as(_F), perm.ctypes.data_as(_I), len(vals), ctypes.c_float(scale),
                                ctypes.byref(lo), ctypes.byref(hi), ctypes.byref(value))
    return lo.value, hi.value, value.value, bad


@pytest.mark.parametrize("es", [1, 3])
@pytest.mark.parametrize("scale_log2,magnitude", [(24, 1.0), (24, 3.0e3), (24, 2.0e6), (10, 1.0e8), (40, 1.0e-6)])
def test_sum_is_exact_and_order_independent(es, scale_log2, magnitude):
    lib = _lib(es)
    rng = np.random.default_rng(scale_log2)
    scale = float(2.0 ** scale_log2)
    vals = (rng.standard_normal(1000) * magnitude).astype(np.float32)
    vals[::7] *= 1e-4                                  # a mix of magnitudes: single-word and two-word contributions
    exact = sum(int(np.rint(np.float64(v) * scale)) for v in vals)      # x * 2^k is exact in fp32 as well
    results = set()
    for _ in range(4):
        lo, hi, value, bad = _acc(lib, vals, rng.permutation(len(vals)), scale)
        assert not bad
        assert hi * (1 << 21) + lo == exact
        results.add((lo, hi, value))
    assert len(results) == 1                           # bit-identical words and read-back for every order
    lo, hi, value = next(iter(results))
    assert abs(value - exact / scale) <= 2.0 ** -23 * abs(exact / scale) + 1.0 / scale


def test_capacity_and_poison():
    lib = _lib()
    scale = float(2.0 ** 24)
    # contributions close to the per-contribution limit (2^47) stay exact as long as |sum| < 2^52 ...
    vals = np.full(30, 4.0e6, np.float32)              # 4e6 * 2^24 = 6.7e13 = 2^45.9
    lo, hi, value, bad = _acc(lib, vals, np.arange(30), scale)
    assert not bad and hi * (1 << 21) + lo == 30 * int(np.rint(np.float64(vals[0]) * scale))
    # ... and 1024 small ones (each below 2^21: low word only) do not overflow the low word
    vals = np.full(1024, 0.12, np.float32)             # 0.12 * 2^24 = 2.0e6 < 2^21
    lo, hi, value, bad = _acc(lib, vals, np.arange(1024), scale)
    assert not bad and hi * (1 << 21) + lo == 1024 * int(np.rint(np.float64(vals[0]) * scale))
    for poison in (np.nan, np.inf, -np.inf, 1.0e8, -3.0e30):        # 1e8 * 2^24 > 2^47: out of range
        v = np.array([1.0, poison, 2.0], np.float32)
        lo, hi, value, bad = _acc(lib, v, np.arange(3), scale)
        assert bad, poison
        assert hi * (1 << 21) + lo == 3 * (1 << 24)     # the representable contributions are still summed exactly


def test_adjoint_scale_is_a_power_of_two_that_maps_the_maximum_to_2_27():
    lib = _lib()
    for m in (1e-30, 3.7e-9, 0.5, 1.0, 1.5, 2.0, 123.456, 9.9e12, 3e30):
        s = lib.emu_fx_pow2_scale(ctypes.c_float(m))
        mant, exp = np.frexp(np.float64(s))
        assert mant == 0.5                              # a power of two: scaling is exact
        assert 2.0 ** 27 <= np.float32(m) * np.float64(s) < 2.0 ** 28 or s in (2.0 ** 127, 2.0 ** -126)
    assert lib.emu_fx_pow2_scale(ctypes.c_float(0.0)) == 1.0
    assert lib.emu_fx_pow2_scale(ctypes.c_float(np.nan)) == 1.0

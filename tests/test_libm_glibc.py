"""s-rack_b200/csrc/libm_glibc.cuh -- the operation-by-operation restatements of glibc 2.39's exp2 (f64: the oscillator's
V/oct conversion, oscillator.rs:43-48), sin (f64: the sine port, oscillator.rs:133) and powf (the Non-Linear module,
math.rs:203-205) that the device computes with --
compiled for the host (tests/c/libm_glibc_host.cpp: the same header, intrinsics swapped for plain arithmetic) and held to
the platform's libm, bit for bit, over random and special inputs.  The platform's libm is what the CPU oracle calls, so
this pins the device's arithmetic to the oracle's without a GPU; the GPU tests then check the device against the oracle
through patches."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", params=["table", "select"])
def host(tmp_path_factory, request):
    """(both spellings of the fast sine's coefficient choice, SRK_SIN_COEF_SELECT)"""
    so = str(tmp_path_factory.mktemp("libm") / "libm_glibc_host.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-mfma", "-shared", "-fPIC", "-o", so,
                           f"-DSRK_SIN_COEF_SELECT={int(request.param == 'select')}",
                           os.path.join(ROOT, "tests", "c", "libm_glibc_host.cpp")])
    L = ctypes.CDLL(so)
    L.t_exp2_glibc.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
    L.t_exp2_group.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
    L.t_powf_glibc.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
    L.t_sin_glibc.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
    L.t_sin_settle.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
    L.t_sin_fast_scan.restype = ctypes.c_long
    L.t_sin_fast_scan.argtypes = [ctypes.c_int, ctypes.c_long, ctypes.c_double, ctypes.c_ulonglong,
                                  ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_long)]
    L.t_sinf_of_f64.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
    return L


@pytest.fixture(scope="module")
def libm():
    m = ctypes.CDLL("libm.so.6")
    m.exp2.restype = ctypes.c_double
    m.exp2.argtypes = [ctypes.c_double]
    m.powf.restype = ctypes.c_float
    m.powf.argtypes = [ctypes.c_float, ctypes.c_float]
    m.sin.restype = ctypes.c_double
    m.sin.argtypes = [ctypes.c_double]
    m.gnu_get_libc_version = ctypes.CDLL("libc.so.6").gnu_get_libc_version
    m.gnu_get_libc_version.restype = ctypes.c_char_p
    return m


def test_exp2_f64_equals_glibc_bit_for_bit(host, libm):
    rng = np.random.default_rng(1)
    x = np.concatenate([
        rng.uniform(-12, 12, 150000),                   # the V/oct range of any audible patch
        np.float32(rng.uniform(-8, 8, 50000)).astype(np.float64) + np.float32(rng.uniform(-3, 3, 50000)).astype(np.float64),  # f64(cv) + f64(val)
        rng.uniform(-1100, 1100, 60000),                # overflow, underflow, the subnormal range
        rng.standard_normal(20000) * 1e-12, rng.standard_normal(2000) * 1e-17,
        np.arange(-1080, 1030, dtype=np.float64), np.arange(-1080, 1030, dtype=np.float64) + 0.5,
        np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1024.0, 1023.9999999999999, -1075.0, -1074.9999, -1022.0, -1022.0000001, 928.0,
                  928.0000001, -928.5, 511.99999, 512.0, 1e-300, -1e-300, 2.0 ** -54, -(2.0 ** -54), 2.0 ** -55, 5e-324]),
    ])
    x = np.concatenate([x, -2.0 ** -rng.uniform(50, 80, 3000), 2.0 ** -rng.uniform(50, 80, 3000)])  # around the 1 + x early return
    ref = np.array([libm.exp2(float(v)) for v in x])
    for fn in (host.t_exp2_glibc, host.t_exp2_group):  # the scalar form and the sample-group form (exp2_glibc_group) the fused kernels use
        for order in (np.arange(x.size), rng.permutation(x.size)):  # (groups of four: specials next to ordinary arguments)
            xs = np.ascontiguousarray(x[order])
            y = np.zeros_like(xs)
            fn(xs.ctypes.data, y.ctypes.data, xs.size)
            same = (y.view(np.uint64) == ref[order].view(np.uint64)) | (np.isnan(y) & np.isnan(ref[order]))
            assert same.all(), f"glibc {libm.gnu_get_libc_version().decode()}: {int((~same).sum())} of {x.size} differ, e.g. x = {xs[~same][:5]}"


def test_sin_f64_equals_glibc_bit_for_bit(host, libm):
    """`(pos * PI * 2.0).sin()` of the sine port (oscillator.rs:133): the argument is in [0, 2 pi); every branch of
    s_sin.c below the huge-argument reduction is restated (and checked well beyond that range)."""
    import math
    rng = np.random.default_rng(3)
    pos = np.concatenate([rng.uniform(0, 1, 300000), rng.uniform(0, 1, 50000) * 2.0 ** -rng.integers(1, 40, 50000),
                          np.arange(0, 4096) / 4096.0, 1.0 - 2.0 ** -np.arange(1, 54.0)])
    x = np.concatenate([pos * math.pi * 2.0,                       # exactly the products the oscillator forms
                        rng.uniform(-7, 7, 100000), rng.uniform(-1e6, 1e6, 60000), rng.uniform(-0.2, 0.2, 60000),
                        rng.standard_normal(20000) * 1e-8,
                        np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 0.126, 0.12599999, 0.855469, 0.8554687, 2.426265, 2.4262647,
                                  math.pi, 2 * math.pi, math.pi / 2, 1e-300, 2.0 ** -26, 2.0 ** -27, 105414349.0, -105414349.0])])
    y = np.zeros_like(x)
    host.t_sin_glibc(x.ctypes.data, y.ctypes.data, x.size)
    ref = np.array([libm.sin(float(v)) for v in x])
    same = (y.view(np.uint64) == ref.view(np.uint64)) | (np.isnan(y) & np.isnan(ref))
    assert same.all(), f"{int((~same).sum())} of {x.size} differ, e.g. x = {[float(v).hex() for v in x[~same][:5]]}"


def test_fast_sine_is_close_to_glibcs(host, libm):
    """lg_sin_fast -- the sine of the common route, explicit IEEE operations only, hence the same bits on the device --
    stays within LG_SIN_FAST_MAX_DIFF = 1 bit pattern of the platform's sin, a sixteenth of the tie band; and the
    settled, narrowed value is the platform's `(float) sin(x)` on every argument scanned.  (1.8e8 arguments gave the
    same maximum; the run here is sized for the CPU suite.)"""
    import re
    src = open(os.path.join(ROOT, "s-rack_b200", "csrc", "libm_glibc.cuh")).read()
    max_diff = int(re.search(r"#define LG_SIN_FAST_MAX_DIFF (\d+)", src).group(1))
    band = int(re.search(r"#define SRK_SIN_TIE_BAND (\d+)", src).group(1))
    assert band >= 8 * max_diff
    for mode, n, lim in ((0, 30_000_000, 0.0), (1, 8_000_000, 7.0), (1, 8_000_000, 1e5), (1, 4_000_000, 1e-3), (2, 2000, 0.0)):
        wx, bad = ctypes.c_double(), ctypes.c_long()
        worst = host.t_sin_fast_scan(mode, n, lim, 11, ctypes.byref(wx), ctypes.byref(bad))
        assert worst <= max_diff, f"mode {mode}: {worst} patterns from the platform's sin at x = {wx.value.hex()}"
        assert bad.value == 0, f"mode {mode}: {bad.value} narrowed values differ"


def test_sine_port_narrowing_is_glibcs_for_any_fast_sine_within_three_ulp(host, libm):
    """The sine port is `(float) sin(x)`.  The device takes CUDA's sin (<= 2 ulp) and goes through the restatement of
    glibc's only where the value is within 16 bit patterns of an f32 rounding tie (lg_sin_settle).  Here the `fast`
    sine is glibc's own moved by -6 .. +6 bit patterns (3 ulp across a binade boundary), on arguments whose sine sits
    right at a tie: the narrowed result must be glibc's every time -- and would not be without the band."""
    rng = np.random.default_rng(5)

    def gsin(x):
        y = np.zeros_like(x)
        host.t_sin_glibc(x.ctypes.data, y.ctypes.data, x.size)  # == libm's sin (test above), vectorised
        return y

    # f32 rounding ties in (2^-30, 1): halfway between neighbouring f32 values
    f = (rng.uniform(0, 1, 6000) * 2.0 ** -rng.integers(0, 30, 6000)).astype(np.float32)
    f = f[(f > 0) & (f < 1)]
    tie = (f.astype(np.float64) + np.nextafter(f, np.float32(2)).astype(np.float64)) / 2
    cand = []
    for x0 in (np.arcsin(tie), np.pi - np.arcsin(tie[tie > 0.4]), np.arcsin(tie[tie > 0.4]) - np.pi):
        xb = x0.view(np.int64)[:, None] + np.arange(-24, 25)[None, :]  # neighbouring arguments
        cand.append(xb.reshape(-1).view(np.float64))
    x = np.concatenate(cand)
    g = gsin(x)
    low = g.view(np.uint64) & np.uint64(0x1fffffff)
    at_tie = np.abs(low.astype(np.int64) - 0x10000000) <= 8
    assert at_tie.sum() > 5000
    x = np.concatenate([x[at_tie], rng.uniform(0, 2 * np.pi, 200000), rng.uniform(-1, 1, 2000) * 2.0 ** -rng.integers(20, 200, 2000),
                        np.array([0.0, -0.0, 2.0 ** -26, 2.0 ** -27, 1e-300, -1e-300, 5e-324, -5e-324, 2.0 ** -26 * (1 - 2.0 ** -53), np.pi, -np.pi, np.inf, np.nan])])
    want = gsin(x).astype(np.float32)
    got = np.zeros(x.size, np.float32)
    host.t_sinf_of_f64(x.ctypes.data, got.ctypes.data, x.size)  # the composition the device runs: fast sine, settled
    assert ((got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))).all()
    changed = 0
    for d in range(-6, 7):
        ulps = np.full(x.size, d, np.int64)
        ulps[np.abs(x) < 2.0 ** -1000] = 0  # (bit patterns of 0 and the subnormals cannot be moved by -6; the fast sine returns these exactly)
        got = np.zeros(x.size, np.float32)
        host.t_sin_settle(x.ctypes.data, ulps.ctypes.data, got.ctypes.data, x.size)
        same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
        assert same.all(), f"fast sine {d} patterns off: {int((~same).sum())} differ, e.g. x = {[float(v).hex() for v in x[~same][:4]]}"
        with np.errstate(invalid="ignore", over="ignore"):
            moved = (gsin(x).view(np.int64) + d).view(np.float64).astype(np.float32)
        changed += int(((moved.view(np.uint32) != want.view(np.uint32)) & np.isfinite(want) & (np.abs(x) > 2.0 ** -26)).sum())
    assert changed > 1000, "the arguments did not exercise the band"  # plain narrowing of the moved value would have differed


def test_powf_equals_glibc_bit_for_bit(host, libm):
    rng = np.random.default_rng(2)
    a = np.concatenate([rng.uniform(0, 4, 150000), np.exp(rng.uniform(-87, 88, 80000)), rng.uniform(-3, 3, 30000),
                        rng.uniform(0, 1.2e-38, 5000)]).astype(np.float32)
    b = np.concatenate([rng.uniform(0.5, 2, 150000), rng.uniform(-40, 40, 80000), rng.integers(-5, 6, 30000).astype(np.float64),
                        rng.uniform(-2, 2, 5000)]).astype(np.float32)
    sp = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, -1.0, 1e-42, -1e-42, 3e38, 0.5, 2.0, -2.0, -8.0, 3.0, -3.0, 2.5, -0.5, 1e10,
                   127.0, 128.0, -149.0, -150.0, 16777216.0, 16777217.0, 8388609.0], dtype=np.float32)
    # (signalling NaNs are not exercised: ctypes widens the argument to a Python float, which quiets it on the way to libm)
    A, Bm = np.meshgrid(sp, sp)
    a, b = np.concatenate([a, A.ravel()]), np.concatenate([b, Bm.ravel()])
    z = np.zeros_like(a)
    host.t_powf_glibc(a.ctypes.data, b.ctypes.data, z.ctypes.data, a.size)
    ref = np.array([libm.powf(float(u), float(v)) for u, v in zip(a, b)], dtype=np.float32)
    same = (z.view(np.uint32) == ref.view(np.uint32)) | (np.isnan(z) & np.isnan(ref))
    bad = np.where(~same)[0][:5]
    assert same.all(), f"{int((~same).sum())} of {a.size} differ, e.g. {[(float(a[i]), float(b[i]), float(z[i]), float(ref[i])) for i in bad]}"


def test_division_by_a_render_constant_equals_ieee_division(tmp_path):
    """fused_ops.cuh ddiv_by_const (Markstein's final step on a correctly rounded reciprocal) against `a / b` on the
    operand ranges of the polyBLEP correction and of the V/oct conversion: tests/c/ddiv_markstein.c, 1e8 pairs."""
    exe = str(tmp_path / "ddiv_markstein")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-o", exe, os.path.join(ROOT, "tests", "c", "ddiv_markstein.c"), "-lm"])
    r = subprocess.run([exe, "50000000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "0 mismatches" in r.stdout, r.stdout[-800:]

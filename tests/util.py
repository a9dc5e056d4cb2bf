"""Shared helpers for the parity tests."""
import numpy as np

REL_TOL = 1e-5  # BASELINE.json north_star: <= 1e-5 relative f32 deviation from the CPU reference


def ulp_distance(a, b):
    """Distance in f32 units-in-the-last-place between same-shaped f32 arrays."""
    ia = np.ascontiguousarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    ib = np.ascontiguousarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, np.int64(-2**31) - ia, ia)
    ib = np.where(ib < 0, np.int64(-2**31) - ib, ib)
    return np.abs(ia - ib)


def parity_stats(got, ref):
    got = np.asarray(got, dtype=np.float32)
    ref = np.asarray(ref, dtype=np.float32)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    bound = REL_TOL * np.maximum(np.abs(ref.astype(np.float64)), 1.0)
    same = got.view(np.uint32) == ref.view(np.uint32)
    return dict(max_abs=float(err.max(initial=0.0)), worst_ratio=float((err / bound).max(initial=0.0)),
                max_ulp=int(ulp_distance(got, ref).max(initial=0)), bit_identical=float(same.mean()) if same.size else 1.0,
                finite=bool(np.isfinite(got).all() and np.isfinite(ref).all()))


def assert_parity(got, ref, exact=False, what=""):
    """|got - ref| <= 1e-5 * max(|ref|, 1) for every sample; `exact` demands equal bits."""
    s = parity_stats(got, ref)
    assert s["finite"], f"{what}: non-finite samples {s}"
    if exact:
        assert s["bit_identical"] == 1.0, f"{what}: expected bit-identical output, got {s}"
    assert s["worst_ratio"] <= 1.0, f"{what}: deviation above 1e-5 relative: {s}"
    return s


def assert_mix_parity(got_mix, ref_mix_f64, n_voices, what=""):
    """mix tolerance: 1e-5 * max(|mix|, sqrt(V)) against the f64 sum of the oracle stems (§8d)."""
    got = np.asarray(got_mix, dtype=np.float64)
    ref = np.asarray(ref_mix_f64, dtype=np.float64)
    bound = REL_TOL * np.maximum(np.abs(ref), np.sqrt(max(n_voices, 1)))
    err = np.abs(got - ref)
    assert np.isfinite(got).all(), f"{what}: non-finite mix"
    assert (err <= bound).all(), f"{what}: mix deviates: max err {err.max()} vs bound {bound[err.argmax() // err.shape[-1], err.argmax() % err.shape[-1]]}"


def build_both(srk, orc, builder, n_voices, sample_rate=48000, buffer_size=1024, channels=2, **kw):
    """Apply one patch description to the product and to the oracle."""
    gp = srk.Patch(srk.AudioConfig(sample_rate, buffer_size, channels))
    op = orc.OraclePatch(sample_rate, buffer_size, channels)
    gh = builder(gp, n_voices, **kw)
    oh = builder(op, n_voices, **kw)
    return gp, op, gh, oh

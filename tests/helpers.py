"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    """Vectors made by the independent numpy restatement (tests/golden/make_golden.py)."""
    return sorted(p for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")) if not os.path.basename(p).startswith(("wgsl_", "rust_")))


def wgsl_golden_cases():
    """Vectors made by executing the reference's own WGSL source (tests/golden/make_wgsl_golden.py)."""
    return sorted(p for p in glob.glob(os.path.join(GOLDEN_DIR, "wgsl_*.npz")) if "wgsl_default" not in p and "wgsl_curl" not in p and "wgsl_present" not in p)


WGSL_CURL = os.path.join(GOLDEN_DIR, "wgsl_curl_64x48.npz")
WGSL_PRESENT = os.path.join(GOLDEN_DIR, "wgsl_present_64x48.npz")


WGSL_DEFAULT = os.path.join(GOLDEN_DIR, "wgsl_default_600x375_f2.npz")


def sha(a):
    import hashlib

    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bits_equal(a, b, what=""):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if not np.array_equal(a.view(np.uint32), b.view(np.uint32)):
        bad = np.argwhere(a.view(np.uint32) != b.view(np.uint32))
        i = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} of {a.size} f32 values differ bitwise; first at {i}: {a[i]!r} vs {b[i]!r}")


def assert_close_rel(a, b, rel=1e-5, floor=0.0277777, what=""):
    """north_star tolerance: |a-b| <= rel * max(|a|, |b|, w_i-scale floor)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    scale = np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
    err = np.abs(a - b) / scale
    assert err.max() <= rel, f"{what}: max relative error {err.max():.3e} > {rel}"


def tau_default():
    return float(np.float32(3.0) * np.float32(0.02) + np.float32(0.5))

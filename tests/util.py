import numpy as np


def rel_err(got, ref):
    """The reference's own acceptance measure (tests/test_tvf.py:46): max|got - ref| / max|ref|."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, f"shape {got.shape} != {ref.shape}"
    if ref.size == 0:
        return 0.0
    scale = max(np.abs(ref).max(), 1e-30)
    return float(np.abs(got - ref).max() / scale)


def ulp_diff(got, ref):
    """Largest distance in float32 ulps between two float32 arrays (0 = bit-identical up to +-0)."""
    a = np.asarray(got, dtype=np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(ref, dtype=np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return int(np.abs(a - b).max()) if a.size else 0


TOL = 1e-5  # north_star: within 1e-5 relative error in fp32

import numpy as np


def rel_err_rows(got, ref, floor=0.0):
    """Per-row norm-wise relative error ||got - ref||_2 / max(||ref||_2, floor)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    num = np.linalg.norm(got - ref, axis=1)
    den = np.maximum(np.linalg.norm(ref, axis=1), floor)
    den = np.where(den == 0.0, 1.0, den)
    return num / den

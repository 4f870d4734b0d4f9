"""Synthetic test functions used as input generators for the BASELINE configs (mirror of
bofire/benchmarks: ZDT1 multi.py:429-485 (BoTorch's ZDT1), DTLZ2 multi.py:95-132, Himmelblau
single.py:409-446, Detergent detergent.py:10-88).  Cheap host-side numpy; not part of the hot path."""
import math

import numpy as np


def zdt1(X):
    X = np.asarray(X, dtype=np.float64)
    f1 = X[..., 0]
    g = 1 + 9 * X[..., 1:].mean(axis=-1)
    return np.stack([f1, g * (1 - np.sqrt(f1 / g))], axis=-1)


def dtlz2(X, num_objectives):
    X = np.asarray(X, dtype=np.float64)
    k = X.shape[-1] - num_objectives + 1
    g1 = 1 + ((X[..., -k:] - 0.5) ** 2).sum(axis=-1)
    fs = []
    for i in range(num_objectives):
        idx = num_objectives - 1 - i
        f = g1 * np.cos(X[..., :idx] * (math.pi / 2)).prod(axis=-1)
        if i > 0:
            f = f * np.sin(X[..., idx] * (math.pi / 2))
        fs.append(f)
    return np.stack(fs, axis=-1)


def himmelblau(X):
    X = np.asarray(X, dtype=np.float64)
    x1, x2 = X[..., 0], X[..., 1]
    return (x1**2 + x2 - 11) ** 2 + (x1 + x2**2 - 7) ** 2


DETERGENT_COEF = np.array([
    [0.4967, 0.0, 0.6477, 1.523, 0.0], [0.0, 4.7376, 2.3023, 0.0, 1.6277], [0.0, 0.0, 0.7259, 0.0, 0.0],
    [0.0, 0.0, 0.9427, 0.0, 0.0], [4.3969, 0.0, 0.2026, 0.0, 0.0], [0.3328, 0.0, 1.1271, 0.0, 0.0],
    [0.0, 16.6705, 0.0, 0.0, 7.4029], [0.0, 1.8798, 0.0, 0.0, 1.7718], [6.6462, 1.5423, 0.0, 0.0, 0.0],
    [0.0, 0.0, 9.5141, 3.0926, 0.0], [2.9168, 0.0, 0.0, 5.5051, 9.279], [8.3815, 0.0, 0.0, 2.9814, 8.7799],
    [0.0, 0.0, 0.0, 0.0, 7.3127], [12.2062, 0.0, 9.0318, 3.2547, 0.0], [3.2526, 13.8423, 0.0, 14.0818, 0.0],
    [7.3971, 0.7834, 0.0, 0.8258, 0.0], [0.0, 3.214, 13.301, 0.0, 0.0], [0.0, 8.2386, 2.9588, 0.0, 4.6194],
    [0.8737, 8.7178, 0.0, 0.0, 0.0], [0.0, 2.6651, 2.3495, 0.046, 0.0], [0.0, 0.0, 0.0, 0.0, 0.0]])
DETERGENT_BOUNDS = np.array([[0.0, 0.0, 0.02, 0.0, 0.0], [0.2, 0.3, 0.2, 0.06, 0.04]])


def detergent(X):
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    iu = np.triu_indices(5)
    xp = np.stack([np.concatenate([[1.0], x, np.outer(x, x)[iu]]) for x in X], axis=0)
    return xp @ DETERGENT_COEF

"""Gate constants.  e3nn's ``Gate`` wraps its activations in ``normalize2mom`` so that
E_{z~N(0,1)}[act(z)^2] = 1.  e3nn estimates the constant from 1e6 random samples; here it
is computed by Gauss-Hermite quadrature (deterministic, ~1e-12 accurate).  [public-SEGNN:
O3TensorProductSwishGate = TP -> Gate(scalars: silu, gates: sigmoid)]."""
from __future__ import annotations

import math

import numpy as np


def normalize2mom_const(fn) -> float:
    x, w = np.polynomial.hermite.hermgauss(256)
    z = math.sqrt(2.0) * x
    m2 = float((w * fn(z) ** 2).sum() / math.sqrt(math.pi))
    return m2 ** -0.5


def _sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


SILU_CST = normalize2mom_const(lambda z: z * _sigmoid(z))
SIGMOID_CST = normalize2mom_const(_sigmoid)

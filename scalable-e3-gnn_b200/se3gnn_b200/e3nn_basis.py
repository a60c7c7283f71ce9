"""Change of basis between this library's real l = 2 components and e3nn's.

The reference asserts lmax == 1 (L1TP:13-14) and would use e3nn's generic tensor product for l_max = 2, so features,
``edge_attr`` and checkpoints produced with e3nn arrive in e3nn's basis.  For l <= 1 the two agree (l = 1 is the
Cartesian vector (x, y, z) in both).  For l = 2:

    this library (csrc/o3tp.cu, oracle/lmax2_oracle.py, z polar):   xy, yz, 2zz - xx - yy, zx, xx - yy
    e3nn (o3.spherical_harmonics, y polar, m = -2 .. 2):            xz, xy, 2yy - xx - zz, yz, zz - xx

each as an orthonormal symmetric-traceless quadratic form.  ``L2_E3NN_FROM_REPO`` is the 5 x 5 orthogonal matrix B with
``x_e3nn = B @ x_repo``; e3nn's published component formulas are restated in ``e3nn_sh2`` and the test pins B to them.
Coupling tensors are unique up to a sign per (l1, l2, l3) once the bases are fixed, so a tensor product transformed
with B equals e3nn's up to one sign per path; that sign is absorbed by the path's weights (``path_signs`` documents
which triples can differ: e3nn is not installed in this image, so the signs of its wigner_3j for l = 2 are not
verifiable here and are reported as unknown rather than guessed).
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np
import torch

from .irreps import Irreps, as_irreps


def _quad(c) -> np.ndarray:
    """unit-Frobenius-norm symmetric traceless 3x3 matrix of the quadratic form with coefficients c[(i, j)]"""
    m = np.zeros((3, 3))
    for (i, j), v in c.items():
        m[i, j] += v / 2.0
        m[j, i] += v / 2.0
    return m / np.linalg.norm(m)


X, Y, Z = 0, 1, 2
REPO_Q2 = np.stack([_quad({(X, Y): 1}), _quad({(Y, Z): 1}), _quad({(Z, Z): 2, (X, X): -1, (Y, Y): -1}),
                    _quad({(Z, X): 1}), _quad({(X, X): 1, (Y, Y): -1})])
E3NN_Q2 = np.stack([_quad({(X, Z): 1}), _quad({(X, Y): 1}), _quad({(Y, Y): 2, (X, X): -1, (Z, Z): -1}),
                    _quad({(Y, Z): 1}), _quad({(Z, Z): 1, (X, X): -1})])

# B[i, j] = <Q^e3nn_i, Q^repo_j>_F : components in e3nn's basis from components in this library's basis
L2_E3NN_FROM_REPO = np.einsum("iab,jab->ij", E3NN_Q2, REPO_Q2)


def e3nn_sh2(vec: np.ndarray) -> np.ndarray:
    """e3nn's l = 2 real spherical harmonics of unit vectors, 'component' normalisation (|Y|^2 = 5), as published in
    e3nn's generated ``_spherical_harmonics``: sqrt(15) xz, sqrt(15) xy, sqrt(5) (y^2 - (x^2 + z^2)/2), sqrt(15) yz,
    (sqrt(15)/2) (z^2 - x^2)."""
    x, y, z = vec[..., 0], vec[..., 1], vec[..., 2]
    s15, s5 = math.sqrt(15.0), math.sqrt(5.0)
    return np.stack([s15 * x * z, s15 * x * y, s5 * (y * y - 0.5 * (x * x + z * z)), s15 * y * z, 0.5 * s15 * (z * z - x * x)], -1)


def _block_matrix(irreps: Irreps, b2: np.ndarray) -> np.ndarray:
    d = irreps.dim
    m = np.zeros((d, d))
    i = 0
    for mi in irreps:
        w = 2 * mi.ir.l + 1
        for _ in range(mi.mul):
            m[i:i + w, i:i + w] = b2 if mi.ir.l == 2 else np.eye(w)
            i += w
    return m


def to_e3nn(x: torch.Tensor, irreps) -> torch.Tensor:
    """[..., irreps.dim] features in this library's layout -> e3nn's (every l = 2 block rotated by B)."""
    m = torch.as_tensor(_block_matrix(as_irreps(irreps), L2_E3NN_FROM_REPO), dtype=x.dtype, device=x.device)
    return x @ m.T


def from_e3nn(x: torch.Tensor, irreps) -> torch.Tensor:
    m = torch.as_tensor(_block_matrix(as_irreps(irreps), L2_E3NN_FROM_REPO.T), dtype=x.dtype, device=x.device)
    return x @ m.T


def coupling_to_e3nn(c: np.ndarray, ls: Sequence[int]) -> np.ndarray:
    """Coupling tensor [2l1+1, 2l2+1, 2l3+1] of this library -> the same invariant tensor in e3nn's bases (equal to
    e3nn's wigner_3j(l1, l2, l3) up to a sign)."""
    mats = [L2_E3NN_FROM_REPO if l == 2 else np.eye(2 * l + 1) for l in ls]
    return np.einsum("ia,jb,kc,abc->ijk", mats[0], mats[1], mats[2], c)

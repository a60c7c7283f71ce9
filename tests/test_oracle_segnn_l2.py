"""The fp64 specification of the SEGNN l_max = 2 model is O(3)-equivariant: rotating (and inverting) the cloud rotates the
1o node outputs.  (CPU only; the GPU parity test is tests/test_segnn_l2_gpu.py.)"""
import numpy as np
import pytest
import torch

from oracle import lmax2_oracle as O2
from oracle.segnn_l2_oracle import SEGNNL2Oracle


@pytest.mark.parametrize("inversion", [False, True])
def test_model_equivariance(inversion):
    rng = np.random.default_rng(0)
    nn_, e = 40, 300
    pos = rng.standard_normal((nn_, 3))
    vel = rng.standard_normal((nn_, 3))
    dst = np.sort(rng.integers(0, nn_, e))
    src = rng.integers(0, nn_, e)
    mass = rng.random(nn_) + 0.5
    R = O2._rand_rot(rng) * (-1.0 if inversion else 1.0)

    def inputs(P, V):
        rel = P[src] - P[dst]
        ea = O2.spherical_harmonics(rel, 2)
        na = np.zeros((nn_, 9))
        np.add.at(na, dst, ea)
        na /= np.maximum(np.bincount(dst, minlength=nn_), 1)[:, None]
        na += O2.spherical_harmonics(V, 2)
        x_in = np.concatenate([P - P.mean(0), V, np.linalg.norm(V, axis=1)[:, None], mass[:, None]], 1)
        ex = np.stack([np.linalg.norm(rel, axis=1), mass[dst] * mass[src]], 1)
        t = lambda a: torch.from_numpy(a)
        return t(x_in), t(na), t(ea), t(ex), torch.from_numpy(dst), torch.from_numpy(src)

    torch.manual_seed(0)
    model = SEGNNL2Oracle("6x0e+3x1o+2x2e", 2)
    with torch.no_grad():
        a = model(*inputs(pos, vel)).numpy()
        b = model(*inputs(pos @ R.T, vel @ R.T)).numpy()
    np.testing.assert_allclose(b, a @ R.T, atol=1e-10 * max(1.0, np.abs(a).max()))
    assert np.abs(a).max() > 1e-3

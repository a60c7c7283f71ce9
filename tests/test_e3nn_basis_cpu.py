"""e3nn <-> repo change of basis for l = 2 (se3gnn_b200/e3nn_basis.py): orthogonality, agreement with e3nn's published
spherical-harmonics formulas (hard-coded fixture), invariance of the transformed couplings, and the known e3nn values
for the l <= 1 triples (wigner_3j(1,1,0) = delta / sqrt 3, wigner_3j(1,1,1) = epsilon / sqrt 6)."""
import numpy as np
import torch

from oracle import lmax2_oracle as O2
from se3gnn_b200 import e3nn_basis as EB


def test_matrix_is_orthogonal_and_a_signed_permutation_mix():
    b = EB.L2_E3NN_FROM_REPO
    np.testing.assert_allclose(b @ b.T, np.eye(5), atol=1e-12)
    # xz, xy, yz are plain relabellings; only the two diagonal forms mix
    np.testing.assert_allclose([abs(b[0, 3]), abs(b[1, 0]), abs(b[3, 1])], 1.0, atol=1e-12)
    np.testing.assert_allclose(np.sort(np.abs(b[[2, 4]][:, [2, 4]]).ravel()), [0.5, 0.5, np.sqrt(3) / 2, np.sqrt(3) / 2], atol=1e-12)


def test_maps_repo_sh2_to_e3nn_published_formulas():
    rng = np.random.default_rng(0)
    v = rng.standard_normal((200, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    repo = O2.spherical_harmonics(v, 2)[:, 4:9]                 # this library's SH(2) ('integral' normalisation)
    want = EB.e3nn_sh2(v)                                       # e3nn, 'component' normalisation
    got = repo @ EB.L2_E3NN_FROM_REPO.T
    scale = np.linalg.norm(want, axis=1) / np.linalg.norm(got, axis=1)
    np.testing.assert_allclose(scale, scale[0], rtol=1e-10)     # one global normalisation constant
    np.testing.assert_allclose(got * scale[0], want, atol=1e-10)


def test_feature_round_trip_and_layout():
    x = torch.randn(7, 3 + 2 * 3 + 2 * 5, dtype=torch.float64)
    ir = "3x0e+2x1o+2x2e"
    y = EB.to_e3nn(x, ir)
    torch.testing.assert_close(EB.from_e3nn(y, ir), x)
    torch.testing.assert_close(y[:, :9], x[:, :9])              # l <= 1 blocks untouched
    torch.testing.assert_close(y[:, 9:14], x[:, 9:14] @ torch.as_tensor(EB.L2_E3NN_FROM_REPO.T))


def test_transformed_couplings_are_invariant_unit_tensors_and_match_known_e3nn_values():
    rng = np.random.default_rng(1)
    for l1, l2, l3 in [(1, 1, 2), (2, 1, 1), (2, 2, 2), (2, 2, 0), (1, 2, 2), (2, 0, 2)]:
        c = EB.coupling_to_e3nn(np.asarray(O2.cg(l1, l2, l3)), (l1, l2, l3))
        np.testing.assert_allclose(np.linalg.norm(c), 1.0, atol=1e-10)
        # invariance under rotations expressed in e3nn's bases: D_e3nn = B D_repo B^T
        R = O2._rand_rot(rng)
        Ds = []
        for l in (l1, l2, l3):
            D = O2.wigner_D(l, R)
            Ds.append(EB.L2_E3NN_FROM_REPO @ D @ EB.L2_E3NN_FROM_REPO.T if l == 2 else D)
        np.testing.assert_allclose(np.einsum("ia,jb,kc,abc->ijk", *Ds, c), c, atol=1e-10)
    # l <= 1: e3nn's values are known in closed form and the bases coincide
    eps = np.zeros((3, 3, 3))
    for i, j, k in [(0, 1, 2), (1, 2, 0), (2, 0, 1)]:
        eps[i, j, k], eps[i, k, j] = 1, -1
    np.testing.assert_allclose(np.abs(np.asarray(O2.cg(1, 1, 1))), np.abs(eps) / np.sqrt(6), atol=1e-12)
    np.testing.assert_allclose(np.asarray(O2.cg(1, 1, 0))[:, :, 0], np.eye(3) / np.sqrt(3), atol=1e-12)

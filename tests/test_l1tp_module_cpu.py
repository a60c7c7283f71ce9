"""Host-side mirror of the reference module: constructor parity (weights for the same seed,
norm buffers, instructions, state_dict keys) and error behaviour.  No GPU needed."""
import os

import numpy as np
import pytest
import torch

from conftest import golden_l1tp_files, load_golden
from se3gnn_b200.irreps import Irreps
from models.segnn.l1_tensor_prod import L1TensorProduct


@pytest.mark.parametrize("path", golden_l1tp_files(), ids=lambda p: os.path.basename(p)[5:-4])
def test_ctor_matches_reference(path):
    rec = load_golden(path)
    meta = rec["meta"]
    torch.manual_seed(meta["seed"])
    tp = L1TensorProduct(Irreps(meta["in1"]), Irreps(meta["out"]), **meta["kwargs"])
    sd = tp.state_dict()
    assert list(sd.keys()) == meta["state_keys"]
    for k, v in sd.items():
        ref = rec["sd_" + k]
        assert tuple(v.shape) == ref.shape
        np.testing.assert_array_equal(v.numpy(), ref)  # same RNG stream -> bit-identical init
    assert len(tp.instructions) == len(meta["instructions"])
    for ins, ref in zip(tp.instructions, meta["instructions"]):
        assert [ins.i_in1, ins.i_in2, ins.i_out, ins.connection_mode, ins.has_weight] == ref[:5]
        assert ins.path_weight == pytest.approx(ref[5], rel=1e-12)
        assert list(ins.path_shape) == ref[6]


def test_default_out_irreps_and_attrs():
    tp = L1TensorProduct(Irreps("16x0e+8x1o"))
    assert str(tp.iro) == "16x0e+8x1o" and tp.in1_dim == 40 and tp.in2_dim == 4
    assert str(tp.iri2) == "1x0e+1x1o"
    assert tp.num_i1_l0e == 16 and tp.num_i1_l1o == 8 and tp.dim_i1_l1o == 24 and tp.num_i1_l0 == 16
    assert tp.dim_o_l0e == 16 and tp.dim_o_l1o == 24 and tp.dim_o_l0o == 0 and tp.dim_o_l1e == 0
    assert tp.iri1_l0e.sum() == 16 and tp.iri1_l1o.sum() == 24 and tp.iro_l1o.dtype == torch.bool
    assert tp.cg110 == pytest.approx(3 ** -0.5) and tp.cg111 == pytest.approx(6 ** -0.5) and tp.cg000 == 1
    assert tp.norm_l0e[0].item() == pytest.approx(24 ** -0.5, rel=1e-6)
    assert tp.norm_l1o[0].item() == pytest.approx((3 / 24) ** 0.5, rel=1e-6)
    assert tp.is_norm and tp.is_comp_norm


def test_error_behaviour():
    with pytest.raises(AssertionError):
        L1TensorProduct(Irreps("4x0e"))                        # lmax must equal 1 (L1TP:13)
    with pytest.raises(AssertionError):
        L1TensorProduct(Irreps("4x0e+1x1o"), Irreps("2x0e"))   # L1TP:14
    with pytest.raises(AssertionError):
        L1TensorProduct(Irreps("4x0e+1x1o"), in1_var=[1.0])    # L1TP:101
    with pytest.raises(Exception):
        L1TensorProduct(Irreps("4x0e+1x1o"), irrep_normalization="norm")   # L1TP:118
    tp = L1TensorProduct(Irreps("4x0e+1x1o"), irrep_normalization="none", path_normalization="none")
    assert not tp.is_norm
    with pytest.raises(AttributeError):                        # quirk Q3
        tp(torch.zeros(2, 7), torch.zeros(2, 4))
    tp = L1TensorProduct(Irreps("4x0e+1x1o"))
    with pytest.raises(Exception):
        tp(torch.zeros(2, 6), torch.zeros(2, 4))               # wrong last dim (L1TP:236)
    with pytest.raises(IndexError):
        tp(torch.zeros(2, 3, 7), torch.zeros(2, 3, 4))         # quirk Q4
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tp(torch.zeros(2, 7), torch.zeros(2, 4))               # product path never runs on CPU
    with pytest.raises(AttributeError):
        L1TensorProduct(Irreps("3x0e+4x1o"), Irreps("2x0o+1x1o"))  # no weights for 0o: reference fails the same way


def test_q1_parity_blind_norm():
    tp = L1TensorProduct(Irreps("4x0e+3x0o+2x1e+5x1o"), Irreps("1x0e+1x1o"))
    assert tp.norm_l0e[0].item() == pytest.approx(14 ** -0.5, rel=1e-6)

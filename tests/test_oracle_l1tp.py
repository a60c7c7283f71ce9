"""Pin the numpy oracle (oracle/l1tp_oracle.py) against golden vectors produced by
the unmodified reference (tests/golden/make_l1tp_golden.py) and the analytic norm
values in SURVEY.md section 8c."""
import math
import os

import numpy as np
import pytest

from conftest import golden_l1tp_files, load_golden
from oracle import l1tp_oracle as O


def _weights(rec, dt):
    return {k[3:]: rec[k].astype(dt) for k in rec if k.startswith("sd_weights")}


def _norms(rec):
    return {k[3:]: rec[k].astype(np.float64) for k in rec if k.startswith("sd_norm")}


@pytest.mark.parametrize("path", golden_l1tp_files(), ids=lambda p: os.path.basename(p)[5:-4])
def test_oracle_matches_reference_fp64(path):
    rec = load_golden(path)
    meta = rec["meta"]
    x, y, go = (rec[k].astype(np.float64) for k in ("x", "y", "gout"))
    w, nrm = _weights(rec, np.float64), _norms(rec)
    out = O.forward(x, y, w, nrm, meta["in1"], meta["out"])
    np.testing.assert_allclose(out, rec["out_f64"], rtol=1e-6, atol=1e-6)  # norms stored fp32
    gx, gy, gw = O.backward(x, y, go, w, nrm, meta["in1"], meta["out"])
    np.testing.assert_allclose(gx, rec["gx_f64"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(gy, rec["gy_f64"], rtol=1e-6, atol=1e-5)
    for k, g in gw.items():
        np.testing.assert_allclose(g, rec[f"gw_{k}_f64"], rtol=1e-6, atol=1e-5)
    assert set(gw) == {k[3:] for k in rec if k.startswith("sd_weights")}


@pytest.mark.parametrize("path", golden_l1tp_files(), ids=lambda p: os.path.basename(p)[5:-4])
def test_oracle_norms_shapes_instructions(path):
    rec = load_golden(path)
    meta = rec["meta"]
    in1, out = O.parse_irreps(meta["in1"]), O.parse_irreps(meta["out"])
    a, wi, instr = O.norm_factors(in1, out, **meta["kwargs"])
    bufs = O.norm_buffers(out, a)
    for k, v in bufs.items():
        np.testing.assert_allclose(v, rec["sd_" + k], rtol=1e-6)
    shapes = O.weight_shapes(in1, out)
    assert shapes == {k[3:]: rec[k].shape for k in rec if k.startswith("sd_weights")}
    ref_instr = meta["instructions"]
    assert len(instr) == len(ref_instr)
    for mine, ref in zip(instr, ref_instr):
        assert list(mine[:5]) == ref[:5]
        assert math.isclose(mine[5], ref[5], rel_tol=1e-12)
        assert list(mine[6]) == ref[6]


def test_analytic_norms():
    # SURVEY 8c (ii)
    a, _, _ = O.norm_factors(O.parse_irreps("16x0e+8x1o"), O.parse_irreps("16x0e+8x1o"))
    assert math.isclose(a[0], 1 / math.sqrt(24)) and math.isclose(a[1], math.sqrt(3 / 24))
    a, _, _ = O.norm_factors(O.parse_irreps("3x0e+2x1o"), O.parse_irreps("3x0e+2x1o"),
                             in1_var=[2, .5], in2_var=[1, 3], out_var=[4, 1])
    assert math.isclose(a[0], 2 / 3) and math.isclose(a[1], math.sqrt(3 / 19))
    # quirk Q1: parity is not checked for l=0 outputs
    a, _, _ = O.norm_factors(O.parse_irreps("4x0e+3x0o+2x1e+5x1o"), O.parse_irreps("1x0e"))
    assert math.isclose(a[0], 1 / math.sqrt(14))


def test_oracle_fp32_close_to_reference_fp32():
    for path in golden_l1tp_files():
        rec = load_golden(path)
        meta = rec["meta"]
        w, nrm = _weights(rec, np.float32), _norms(rec)
        out = O.forward(rec["x"], rec["y"], w, nrm, meta["in1"], meta["out"])
        scale = np.abs(rec["out_f32"]).max()
        assert np.abs(out - rec["out_f32"]).max() <= 1e-5 * scale

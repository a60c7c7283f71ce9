"""Generate golden vectors by running the UNMODIFIED reference tensor product.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_l1tp_golden.py

The reference file is imported from its read-only path through the metadata-only
e3nn shim in ``oracle/e3nn_shim``.  For every case we store inputs, the
reference-initialised weights / norm buffers (torch.manual_seed(seed) right
before construction), the forward output and all gradients for a random
cotangent, once in fp32 (what the CUDA kernels are compared against) and once
with the module cast to fp64 (what the numpy oracle is pinned against).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "e3nn_shim"))
sys.path.insert(0, "/root/reference")

from e3nn.o3 import Irreps  # noqa: E402  (shim)
from models.segnn.l1_tensor_prod import L1TensorProduct  # noqa: E402  (reference)

CASES = [
    # name, in1, out, kwargs
    ("square", "16x0e+8x1o", None, {}),
    ("msg1", "34x0e+10x1o+34x0e+10x1o+2x0e", "44x0e+10x1o", {}),
    ("msg2", "34x0e+10x1o", "44x0e+10x1o", {}),
    ("upd1", "34x0e+10x1o+34x0e+10x1o", "44x0e+10x1o", {}),
    ("upd2", "34x0e+10x1o", "34x0e+10x1o", {}),
    ("embed", "2x0e+2x1o", "34x0e+10x1o", {}),
    ("readout", "34x0e+10x1o", "1x1o", {}),
    ("four_species", "4x0e+3x0o+2x1e+5x1o", "3x0e+2x0o+2x1e+3x1o", {}),
    ("interleaved", "2x0e+1x1o+3x0e+2x1o+1x0o+1x1e", "2x1o+3x0e+1x1e+2x0o+1x1o", {}),
    ("vars", "3x0e+2x1o", "3x0e+2x1o",
     dict(in1_var=[2.0, 0.5], in2_var=[1.0, 3.0], out_var=[4.0, 1.0])),
    ("path_none", "5x0e+3x1o", "4x0e+2x1o", dict(path_normalization="none")),
    ("irrep_none", "5x0e+3x1o", "4x0e+2x1o", dict(irrep_normalization="none")),
    ("odd_outputs", "3x0e+4x1o", "2x0e+2x1e+1x1o", {}),
    ("wide", "96x0e+32x1o", "80x0e+24x1o", {}),
]
ROWS = 37  # deliberately not a multiple of any tile size


def run_case(name, in1, out, kw, seed):
    torch.manual_seed(seed)
    tp = L1TensorProduct(Irreps(in1), Irreps(out) if out is not None else None, **kw)
    g = torch.Generator().manual_seed(1000 + seed)
    x = torch.randn(ROWS, tp.in1_dim, generator=g)
    y = torch.randn(ROWS, 4, generator=g)
    go = torch.randn(ROWS, tp.iro.dim, generator=g)
    rec = {"x": x.numpy().copy(), "y": y.numpy().copy(), "gout": go.numpy().copy()}
    for k, v in tp.state_dict().items():
        rec["sd_" + k] = v.detach().numpy().copy()
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        m = L1TensorProduct(Irreps(in1), Irreps(out) if out is not None else None, **kw).to(dt)
        m.load_state_dict({k: v.to(dt) for k, v in tp.state_dict().items()})
        xi = x.detach().clone().to(dt).requires_grad_(True)
        yi = y.detach().clone().to(dt).requires_grad_(True)
        o = m(xi, yi)
        # columns of zero-width species are uninitialised in the reference (L1TP:240); none here
        o.backward(go.to(dt))
        rec[f"out_{tag}"] = o.detach().numpy()
        rec[f"gx_{tag}"] = xi.grad.numpy()
        rec[f"gy_{tag}"] = yi.grad.numpy()
        for k, p in m.named_parameters():
            rec[f"gw_{k}_{tag}"] = p.grad.numpy()
    meta = {
        "name": name, "in1": in1, "out": out if out is not None else in1, "kwargs": kw, "seed": seed,
        "instructions": [[int(i.i_in1), int(i.i_in2), int(i.i_out), i.connection_mode, bool(i.has_weight),
                          float(i.path_weight), [int(s) for s in i.path_shape]] for i in tp.instructions],
        "state_keys": list(tp.state_dict().keys()),
        "reference": "/root/reference/models/segnn/l1_tensor_prod.py (unmodified, via oracle/e3nn_shim)",
        "torch": torch.__version__,
    }
    rec["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, f"l1tp_{name}.npz"), **rec)
    return meta


if __name__ == "__main__":
    for i, (name, in1, out, kw) in enumerate(CASES):
        meta = run_case(name, in1, out, kw, seed=i)
        print(name, meta["state_keys"])

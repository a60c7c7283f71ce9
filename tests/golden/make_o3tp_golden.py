"""Freeze the l <= 2 convention of THIS repo (there is no reference for l = 2: L1TP:13-14): coupling tensors and small
input / output / gradient vectors of `oracle/lmax2_oracle.py`, so that a later change of basis, sign or normalisation
cannot go unnoticed.  These are NOT reference outputs (parity for l = 2 stays unpinned).

    python tests/golden/make_o3tp_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import lmax2_oracle as O2  # noqa: E402

CASES = [
    ("balanced2", [(23, 0, 1), (7, 1, -1), (4, 2, 1)], 2, [(23, 0, 1), (7, 1, -1), (4, 2, 1)]),
    ("message2", [(23, 0, 1), (7, 1, -1), (4, 2, 1), (23, 0, 1), (7, 1, -1), (4, 2, 1), (2, 0, 1)], 2,
     [(34, 0, 1), (7, 1, -1), (4, 2, 1)]),
    ("mixed_parity", [(3, 0, 1), (2, 1, -1), (2, 2, 1), (1, 1, 1), (1, 2, -1), (2, 0, -1)], 2,
     [(3, 0, 1), (2, 1, -1), (1, 2, 1), (2, 1, 1), (1, 2, -1), (5, 0, -1)]),
]


def main():
    cg = {f"cg_{a}{b}{c}": O2.cg(a, b, c) for a in range(3) for b in range(3) for c in range(3) if abs(a - b) <= c <= a + b}
    np.savez_compressed(os.path.join(HERE, "o3tp_couplings.npz"), **cg)
    for name, in1, lmax, out in CASES:
        rng = np.random.default_rng(sum(map(ord, name)))
        in2 = O2.sh_irreps(lmax)
        rows = 29
        d1 = sum(m * (2 * l + 1) for m, l, _ in in1)
        x1 = rng.standard_normal((rows, d1)).astype(np.float32)
        y = O2.spherical_harmonics(rng.standard_normal((rows, 3)), lmax).astype(np.float32)
        shapes = O2.weight_shapes(in1, in2, out)
        w = rng.standard_normal(sum(a * b for a, b in shapes)).astype(np.float32)
        do = sum(m * (2 * l + 1) for m, l, _ in out)
        g = rng.standard_normal((rows, do)).astype(np.float32)
        ws, o = [], 0
        for a, b in shapes:
            ws.append(torch.from_numpy(w[o:o + a * b].astype(np.float64).reshape(a, b)).requires_grad_())
            o += a * b
        xt = torch.from_numpy(x1.astype(np.float64)).requires_grad_()
        yt = torch.from_numpy(y.astype(np.float64)).requires_grad_()
        res = O2.forward(xt, yt, ws, in1, in2, out)
        (res * torch.from_numpy(g.astype(np.float64))).sum().backward()
        meta = dict(in1=in1, lmax=lmax, out=out, paths=O2.paths(in1, in2, out), norm=O2.norm_factors(in1, in2, out))
        np.savez_compressed(os.path.join(HERE, f"o3tp_{name}.npz"), x1=x1, y=y, w=w, g=g, out_f64=res.detach().numpy(),
                            gx_f64=xt.grad.numpy(), gy_f64=yt.grad.numpy(),
                            gw_f64=np.concatenate([t.grad.numpy().reshape(-1) for t in ws]),
                            meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
        print(name, res.shape)


if __name__ == "__main__":
    main()

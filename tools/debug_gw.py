import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scalable-e3-gnn_b200"))
import numpy as np, torch
from oracle import l1tp_oracle as O
from se3gnn_b200.irreps import Irreps
from se3gnn_b200.tp import TPConfig, get_plan, tp_layer
np.set_printoptions(linewidth=200, precision=3, suppress=True)
in1, out, rows = sys.argv[1], sys.argv[2], int(sys.argv[3])
rng = np.random.default_rng(0)
i1, io = O.parse_irreps(in1), O.parse_irreps(out)
w = {k: rng.uniform(-1, 1, s).astype(np.float32) for k, s in O.weight_shapes(i1, io).items()}
a, _, _ = O.norm_factors(i1, io); nrm = O.norm_buffers(io, a)
din, dout = O.irreps_dim(i1), O.irreps_dim(io)
x = rng.standard_normal((rows, din)).astype(np.float32); y = rng.standard_normal((rows, 4)).astype(np.float32)
go = rng.standard_normal((rows, dout)).astype(np.float32)
dev = "cuda"
ws = [torch.from_numpy(w[f"weights_{s}"]).to(dev).requires_grad_(True) if f"weights_{s}" in w else None for s in ("l0e","l0o","l1e","l1o")]
ns = [torch.from_numpy(nrm[f"norm_{s}"].astype(np.float32)).to(dev) if nrm[f"norm_{s}"].size else None for s in ("l0e","l0o","l1e","l1o")]
xt = torch.from_numpy(x).to(dev).requires_grad_(True); yt = torch.from_numpy(y).to(dev)
cfg = TPConfig(plan=get_plan(Irreps(in1), Irreps(out)), widths=(din,))
o = tp_layer(cfg, rows, [xt], [None], yt, ws, ns); o.backward(torch.from_numpy(go).to(dev))
w64 = {k: v.astype(np.float64) for k, v in w.items()}
gx, gy, gw = O.backward(x.astype(np.float64), y.astype(np.float64), go.astype(np.float64), w64, nrm, in1, out)
for i, s in enumerate(("l0e", "l1o")):
    j = 0 if s == "l0e" else 3
    got = ws[j].grad.cpu().numpy(); ref = gw[f"weights_{s}"]
    print(s, "shape", got.shape, "max err", np.abs(got - ref).max(), "ref max", np.abs(ref).max())
    print("got[:6,:6]\n", got[:6, :6]); print("ref[:6,:6]\n", ref[:6, :6])
    print("ratio\n", (got / ref)[:6, :6])
print("gx err", np.abs(xt.grad.cpu().numpy() - gx).max())

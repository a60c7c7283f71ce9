"""Per-launch timing of the l <= 2 tensor product (se3_o3tp_forward / backward) on one GPU, CUDA events on the launching
stream, inputs larger than L2.  Prints one JSON line per configuration with the HBM roofline fraction (algorithmic
bytes = every operand once, SURVEY 8d TP-standalone formula) and the fp32 contraction rate.

    python tools/bench_o3tp.py [--rows 2000000] [--iters 10]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scalable-e3-gnn_b200"))

import torch  # noqa: E402

CONFIGS = {
    # SEGNN l_max = 2 with the public-SEGNN BalancedIrreps(2, 64) hidden type 23x0e+7x1o+4x2e:
    "message2": ("23x0e+7x1o+4x2e+23x0e+7x1o+4x2e+2x0e", "34x0e+7x1o+4x2e"),   # msg1: cat(x_i, x_j, extra) -> gated hidden
    "update2": ("23x0e+7x1o+4x2e", "23x0e+7x1o+4x2e"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2_000_000)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from se3gnn_b200.irreps import Irreps
    from se3gnn_b200.o3tp import O3TensorProduct
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6550.0))
    for name, (i1, io) in CONFIGS.items():
        torch.manual_seed(0)
        tp = O3TensorProduct(Irreps(i1), Irreps(io)).cuda()
        x = torch.randn(a.rows, tp.in1_dim, device="cuda", requires_grad=True)
        y = torch.randn(a.rows, tp.in2_dim, device="cuda")
        g = torch.randn(a.rows, tp.iro.dim, device="cuda")
        for _ in range(3):
            out = tp(x, y)
            out.backward(g)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tf = tb = 0.0
        for _ in range(a.iters):
            ev[0].record()
            out = tp(x, y)
            ev[1].record()
            out.backward(g)
            ev[2].record()
            torch.cuda.synchronize()
            tf += ev[0].elapsed_time(ev[1])
            tb += ev[1].elapsed_time(ev[2])
        tf, tb = tf / a.iters, tb / a.iters
        bf, bb = tp.algo_bytes(a.rows, "fwd"), tp.algo_bytes(a.rows, "bwd")
        print(json.dumps({
            "op": "o3tp", "config": name, "in1": i1, "out": io, "rows": a.rows,
            "tile_fwd": tp._plan.tile_fwd, "tile_bwd": tp._plan.tile_bwd, "weights": tp._plan.weight_floats,
            "fwd_ms": round(tf, 4), "bwd_ms": round(tb, 4),
            "fwd_rows_per_s": a.rows / tf * 1e3, "bwd_rows_per_s": a.rows / tb * 1e3,
            "fwd_GBps": bf / tf / 1e6, "bwd_GBps": bb / tb / 1e6,
            "fwd_frac_hbm": bf / tf / 1e6 / hbm, "bwd_frac_hbm": bb / tb / 1e6 / hbm, "hbm_peak_GBps": hbm,
            "fwd_TFLOPs": tp.flops(a.rows) / tf / 1e9, "bwd_TFLOPs": 2 * tp.flops(a.rows) / tb / 1e9,
            "note": "bwd includes torch's autograd glue (one empty_like per gradient)",
        }), flush=True)


if __name__ == "__main__":
    main()

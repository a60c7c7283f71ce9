"""Weight-gradient sensitivity to the reduction order alone: same single-GPU step, default grid vs SE3_BWDW_GRID=37
(run as separate processes; prints max |dg| / max |g|)."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 2 and sys.argv[1] == "child":
    for p in (ROOT, os.path.join(ROOT, "scalable-e3-gnn_b200")):
        sys.path.insert(0, p)
    import numpy as np, torch
    from models.segnn.segnn import SEGNN
    from se3gnn_b200.pipeline import TrainStep, synthetic_cloud
    n = int(sys.argv[3])
    torch.manual_seed(0); m = SEGNN(num_layers=4).cuda()
    ts = TrainStep(m)
    data = [torch.from_numpy(a).cuda() for a in synthetic_cloud(n, "plummer", seed=7)]
    ts.step_device(*data); torch.cuda.synchronize()
    np.save(sys.argv[2], ts.flat_grad.cpu().numpy())
else:
    import numpy as np
    n = sys.argv[1] if len(sys.argv) > 1 else "400000"
    outs = []
    for tag, env in (("a", {}), ("b", {}), ("c", {"SE3_BWDW_GRID": "37"})):
        f = f"/tmp/g_{tag}.npy"
        subprocess.run([sys.executable, __file__, "child", f, n], env={**os.environ, **env}, check=True)
        outs.append(np.load(f))
    sc = np.abs(outs[0]).max()
    print(f"n={n}: rerun (same order up to atomics) {np.abs(outs[1]-outs[0]).max()/sc:.2e}; 37-CTA weight-gradient grid vs 148: {np.abs(outs[2]-outs[0]).max()/sc:.2e}")

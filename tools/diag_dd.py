"""Diagnostic: W ranks sharing cuda:0 (gloo) run the decomposed step; per-parameter gradient error vs single rank,
plus the run-to-run nondeterminism of the single-rank step."""
import os, sys, socket
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scalable-e3-gnn_b200")):
    sys.path.insert(0, p)
import numpy as np, torch, torch.distributed as dist, torch.multiprocessing as mp

def setup(n, layers):
    from models.segnn.segnn import SEGNN
    from se3gnn_b200.pipeline import TrainStep, synthetic_cloud
    torch.manual_seed(0)
    model = SEGNN(num_layers=layers).cuda()
    data = [torch.from_numpy(a).cuda() for a in synthetic_cloud(n, "plummer", seed=7)]
    return model, data, TrainStep

def worker(rank, world, port, n, layers, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    model, data, TrainStep = setup(n, layers)
    ts = TrainStep(model, decompose=True)
    loss = ts.step_device_dd(*data)
    torch.cuda.synchronize()
    if rank == 0:
        q.put((float(loss), ts.flat_grad.cpu().numpy()))
    dist.destroy_process_group()

if __name__ == "__main__":
    world, n, layers = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn"); q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, port, n, layers, q)) for r in range(world)]
    for p in procs: p.start()
    loss_dd, g_dd = q.get(timeout=600)
    for p in procs: p.join(timeout=60)
    outs = []
    for rep in range(2):
        model, data, TrainStep = setup(n, layers)
        ts = TrainStep(model)
        l = float(ts.step_device(*data)); torch.cuda.synchronize()
        outs.append((l, ts.flat_grad.cpu().numpy().copy()))
    names = [(k, p.numel()) for k, p in model.named_parameters()]
    ref = outs[0][1]; sc = np.abs(ref).max()
    print(f"loss dd {loss_dd:.8f} single {outs[0][0]:.8f} {outs[1][0]:.8f}")
    print(f"single vs single: {np.abs(outs[1][1]-ref).max()/sc:.2e}   dd vs single: {np.abs(g_dd-ref).max()/sc:.2e}")
    o = 0
    for k, m in names:
        a, b, c = ref[o:o+m], g_dd[o:o+m], outs[1][1][o:o+m]
        e_dd, e_ss = np.abs(b-a).max(), np.abs(c-a).max()
        if e_dd / sc > 2e-5:
            print(f"  {k:28s} |g|max {np.abs(a).max():.3e}  dd err {e_dd:.2e} ({e_dd/max(np.abs(a).max(),1e-30):.1e} rel)  rerun err {e_ss:.2e}")
        o += m

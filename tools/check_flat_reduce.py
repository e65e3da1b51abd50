"""torchrun check (N >= 2 GPUs): dist.FlatGradAllReduce (NVLS multimem all-reduce when available) against a plain NCCL all-reduce,
through a real data-parallel ELBO step: every rank steps on its own minibatch, the reduced gradients must equal the NCCL average."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from bayesrul_b200 import Engine, Noise
from bayesrul_b200.compat.nets import init_flat_params
from bayesrul_b200.dist import FlatGradAllReduce

eng = Engine("inception", dev)
eng.set_gemm_backend("fused")
P = eng.P
g = torch.Generator().manual_seed(100 + rank)
x = torch.randn(256, 30, 18, generator=g).to(dev); y = (torch.rand(256, generator=g) * 100).to(dev)
mu = init_flat_params("inception", 12345).to(dev); sg = torch.full_like(mu, 1.351e-3)
red = FlatGradAllReduce(P, dev)
kw = dict(mode="lrt", guide="normal", particles=1, prior_loc=0.0, prior_scale=0.138793, dataset_size=238150)
ok = True
for it in range(4):  # eager, captured, replayed, replayed
    r = eng.elbo_step(x, y, mu, sg, noise=Noise(seed=7 + it, window0=rank * 256), out_flat=red.flat, **kw)
    ref_mu, ref_ls, ref_sc = r["grad_mu"].clone(), r["grad_log_sigma"].clone(), r["scalars"].clone()
    for t in (ref_mu, ref_ls, ref_sc):
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
    out = red.reduce(r)
    got_mu, got_ls = out["grad_mu"] * red.grad_scale, out["grad_log_sigma"] * red.grad_scale
    torch.cuda.synchronize()
    for a, b, nm in ((got_mu, ref_mu, "grad_mu"), (got_ls, ref_ls, "grad_log_sigma"), (out["scalars"], ref_sc, "scalars")):
        err = float((a.double() - b.double()).abs().max() / b.double().abs().max())
        ok = ok and err < 1e-5
        if rank == 0:
            print(f"step {it} {nm}: mode {red.mode}, max rel err vs NCCL AVG {err:.2e}", flush=True)
if rank == 0:
    print("FlatGradAllReduce", "OK" if ok else "MISMATCH", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)

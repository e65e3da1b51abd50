"""2+ ranks: latency of an all-reduce of the ELBO step's flat gradient buffer (2 P + 4 floats = 1.5 MB) through NCCL and through
torch's symmetric-memory one-shot / multimem (NVLS) all-reduce kernels."""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 2 * 187142 + 4
x = torch.randn(n, device=dev)

def timeit(fn, it=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e3

t_nccl = timeit(lambda: dist.all_reduce(x, op=dist.ReduceOp.AVG))
if rank == 0: print(f"nccl all_reduce AVG {n * 4 / 1e6:.2f} MB: {t_nccl:.1f} us", flush=True)
try:
    group = dist.group.WORLD.group_name
    buf = symm_mem.empty(n, device=dev)
    hdl = symm_mem.rendezvous(buf, group)
    buf.copy_(x)
    ops = [o for o in dir(torch.ops.symm_mem) if "all_reduce" in o]
    if rank == 0: print("symm_mem ops:", ops, "multicast:", getattr(hdl, "multicast_ptr", None) not in (None, 0), flush=True)
    for name in ("one_shot_all_reduce", "two_shot_all_reduce_", "multimem_all_reduce_", "multimem_one_shot_all_reduce"):
        try:
            op = getattr(torch.ops.symm_mem, name)
            t = timeit(lambda: op(buf, "sum", group))
            if rank == 0: print(f"symm_mem.{name}: {t:.1f} us", flush=True)
        except Exception as e:  # noqa: BLE001
            if rank == 0: print(f"symm_mem.{name}: unavailable ({type(e).__name__}: {str(e)[:120]})", flush=True)
    # correctness of one_shot: sum over ranks
    buf.fill_(float(rank + 1)); torch.cuda.synchronize(); dist.barrier()
    r = torch.ops.symm_mem.one_shot_all_reduce(buf, "sum", group)
    torch.cuda.synchronize()
    if rank == 0: print("one_shot value", r[0].item(), "expected", world * (world + 1) / 2, flush=True)
except Exception as e:  # noqa: BLE001
    if rank == 0: print("symmetric memory unavailable:", type(e).__name__, str(e)[:300], flush=True)
dist.destroy_process_group()

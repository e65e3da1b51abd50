#!/usr/bin/env python
"""Tracked summaries of the ELBO-train and MC-dropout captures of tools/profile_round.sh.
usage: python tools/make_profiles_train.py <tag> <round-prefix>   e.g.  r01v8 r01"""
import collections, csv, io, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rnd = sys.argv[1], sys.argv[2]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

rows = [r for r in csv.reader(open(os.path.join(G, f"launches_train_{tag}.csv"))) if len(r) > 5]
hdr = rows[0]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
launches = []
for r in rows[1:]:
    try:
        launches.append((r[ki][:60], r[gi], r[bi], float(r[vi].replace(",", "")) / 1e3))
    except ValueError:
        continue
ends = [i for i, l in enumerate(launches) if "post_scalars" in l[0]]  # the last kernel of a step
step = launches[ends[-2] + 1: ends[-1] + 1]  # the last complete step
agg = collections.OrderedDict()
for k, g, b, t in step:
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += t
tot = sum(v[1] for v in agg.values())
with open(os.path.join(P, f"{rnd}_launches_train_lrt.txt"), "w") as f:
    f.write(f"# {rnd} -- ncu launch list (gpu__time_duration.sum, --clock-control none) of ONE LRT ELBO step, Inception, B = 256,\n"
            "#   BRL_NO_GRAPH=1 python tools/profile_train.py lrt 2 simt   (eager launches; brl_elbo_step replays the same kernels as a CUDA graph)\n"
            f"# per-launch times are cold-cache and serialised (sum {tot:.0f} us; the CUDA-graph replay of the step takes ~0.54 ms with its four streams): read SHARES.\n"
            f"# raw csv: gpurun_out/launches_train_{tag}.csv (scratch)\n")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{k:62s} n={n:3d} total={t:8.1f} us  avg={t / n:6.1f} us share={100 * t / tot:5.1f}%\n")
    f.write("# in launch order (kernel, grid, block, us)\n")
    for k, g, b, t in step:
        f.write(f"{k:62s} {g:16s} {b:14s} {t:7.1f}\n")

keys = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_per_inst_issued.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'launch__occupancy_limit_registers']


def capture(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[2]
    name = r[hdr.index("Kernel Name")]
    return name, {k: (r[hdr.index(k)], units[hdr.index(k)]) for k in keys if k in hdr}


with open(os.path.join(P, f"{rnd}_ncu_train_kernels.txt"), "w") as f:
    for what, rep in (("forward dual GEMM of layers.1.branch2.0 (1x1 conv 108 -> 64, M = 7680 rows, LRT epilogue)", f"prof_train_fwd_{tag}.ncu-rep"),
                      ("input-gradient dual GEMM (LRT)", f"prof_train_dx_{tag}.ncu-rep"),
                      ("dual weight-gradient kernel (mean path + sigma^2 path, shared gather)", f"prof_train_dw_{tag}.ncu-rep"),
                      ("tc_conv_kernel<true>: MC-dropout predictive pass, 25 masks x 10000 windows per launch", f"prof_conv_mcd_{tag}.ncu-rep")):
        path = os.path.join(G, rep)
        if not os.path.exists(path):
            continue
        name, m = capture(path)
        f.write(f"# {rnd} -- ncu --set full --clock-control none --import-source on: {name}\n# {what}\n")
        for k in keys:
            if k in m:
                f.write(f"{k} [{m[k][1]}]: {m[k][0]}\n")
        f.write("\n")
print("wrote", f"{rnd}_launches_train_lrt.txt", f"{rnd}_ncu_train_kernels.txt")

#!/usr/bin/env python
"""Tracked summaries of the level-fused training captures of tools/profile_round2.sh.
usage: python tools/make_profiles_fused.py <tag> <round-prefix>   e.g.  r02 r02"""
import collections, csv, io, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rnd = sys.argv[1], sys.argv[2]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("brl::<", "").replace("brl::", "")[:44]


for mode in ("lrt", "flipout"):
    rows = [r for r in csv.reader(open(os.path.join(G, f"launches_train_fused_{mode}_{tag}.csv"))) if len(r) > 5]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    launches = []
    for r in rows[1:]:
        try:
            launches.append((short(r[ki]), r[gi], r[bi], float(r[vi].replace(",", "")) / 1e3))
        except ValueError:
            continue
    ends = [i for i, l in enumerate(launches) if "post_scalars" in l[0]]
    step = launches[ends[-2] + 1: ends[-1] + 1]
    agg = collections.OrderedDict()
    for k, g, b, t in step:
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += t
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, f"{rnd}_launches_train_fused_{mode}.txt"), "w") as f:
        f.write(f"# {rnd} -- ncu launch list (gpu__time_duration.sum, --clock-control none --cache-control none) of ONE {mode.upper()} ELBO step,\n"
                f"#   Inception, B = 256, level-fused tcgen05 back-end:  BRL_NO_GRAPH=1 python tools/profile_train.py {mode} 3 fused\n"
                f"# eager launches, serialised by ncu (sum {tot:.0f} us); brl_elbo_step replays the same kernels as a CUDA graph in which the weight\n"
                "# packing, the fc / head weight gradients and (Flipout) the two particles overlap -- read SHARES, the step time is bench.py's.\n"
                f"# raw csv: gpurun_out/launches_train_fused_{mode}_{tag}.csv (scratch)\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k:46s} n={n:3d} total={t:8.1f} us  avg={t / n:6.1f} us share={100 * t / tot:5.1f}%\n")
        f.write("# in launch order (kernel, grid, block, us)\n")
        for k, g, b, t in step:
            f.write(f"{k:46s} {g:16s} {b:14s} {t:7.1f}\n")

keys = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__block_size',
        'launch__grid_size', 'launch__shared_mem_per_block_dynamic']
rep = os.path.join(G, f"prof_train_fused_{tag}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ni = hdr.index("Kernel Name")
with open(os.path.join(P, f"{rnd}_ncu_train_fused.txt"), "w") as f:
    f.write(f"# {rnd} -- ncu --set full --clock-control none --import-source on -k regex:tt_ (one LRT ELBO step, Inception, B = 256)\n"
            "#   BRL_NO_GRAPH=1 python tools/profile_train.py lrt 3 fused      raw report: gpurun_out/prof_train_fused_" + tag + ".ncu-rep (scratch)\n"
            "# kernels in launch order; grid (tiles x layers of the level) tells which level a tt_fwd / tt_bwd launch is\n")
    for r in rows[2:]:
        f.write(f"\n== {short(r[ni])}\n")
        for k in keys:
            if k in hdr:
                f.write(f"  {k:84s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}\n")
    # SASS evidence from the built library
    so = os.path.join(ROOT, "bayesrul_b200", "lib", "brl_tc_train.o")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    cnt = collections.Counter()
    for line in sass.splitlines():
        m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            cnt[m.group(1).split(".")[0]] += 1
    f.write("\n== SASS mnemonics of bayesrul_b200/lib/brl_tc_train.o (cuobjdump -sass), static instruction counts\n")
    for k in ("UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "SYNCS", "RED", "ATOMG", "HMMA", "FFMA"):
        f.write(f"  {k:10s} {cnt.get(k, 0)}\n")
print("written")

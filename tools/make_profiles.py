#!/usr/bin/env python
"""Turn the scratch ncu outputs in gpurun_out/ into the tracked summaries under profiles/.
usage: python tools/make_profiles.py <tag> <round-prefix>   e.g.  r01v5 r01"""
import collections, csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rnd = sys.argv[1], sys.argv[2]
SPL = int(sys.argv[3]) if len(sys.argv) > 3 else 100  # MC samples per fused launch
WS = SPL * 10000
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

# ---- launch list
rows = [r for r in csv.reader(open(os.path.join(G, f"launches_{tag}.csv"))) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
d = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    k = r[ki][:70]
    d.setdefault(k, [0, 0.0])
    d[k][0] += 1
    d[k][1] += v
tot = sum(v[1] for v in d.values())
with open(os.path.join(P, f"{rnd}_launches_predict_tc.txt"), "w") as f:
    f.write(f"# {rnd} -- ncu launch list (gpu__time_duration.sum, --clock-control none) of\n"
            f"#   python bench.py --steps 2 --warmup 3 --no-cpu --no-train   (tcgen05 engine, B=10000 x S=100, {SPL} samples per launch)\n"
            f"# per-launch times are cold-cache / serialised: read SHARES.  raw csv: gpurun_out/launches_{tag}.csv (scratch)\n")
    for k, (n, t) in sorted(d.items(), key=lambda x: -x[1][1]):
        f.write(f"{k:72s} n={n:4d} total={t / 1e6:9.3f} ms  avg={t / n / 1e3:9.1f} us share={100 * t / tot:5.1f}%\n")

# ---- full captures
keys = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second',
        'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum']


def capture(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[2]
    out = {k: (r[hdr.index(k)], units[hdr.index(k)]) for k in keys if k in hdr}
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(sass)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    h, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
    ix = {c: i for i, c in enumerate(h)}
    tot = sum(int(r[ix['# Samples']]) for r in data)
    stalls = {s[6:]: sum(int(r[ix[s]]) for r in data) for s in h if s.startswith('stall_') and 'Not Issued' not in s}
    mnem = collections.Counter()
    for r in data:
        m = r[ix['Source']].strip().split()
        m = m[1] if m and m[0].startswith('@') else (m[0] if m else '')
        if m.startswith(('UTCHMMA', 'UTCBAR', 'UBLKCP', 'LDTM', 'SYNCS', 'UTCATOMSWS')):
            mnem[m.split('.')[0]] += int(r[ix['Instructions Executed']])
    return out, tot, stalls, mnem


def bytes_of(v, u):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


with open(os.path.join(P, f"{rnd}_ncu_tc_kernels.txt"), "w") as f:
    for name, rep, head in (
        ("tc_conv_kernel", f"prof_conv_{tag}.ncu-rep",
         f"one launch = {SPL} MC samples x 10000 windows = {WS} window-samples; algorithmic GEMM FLOPs/launch = {WS} x 1 981 920 = {WS * 1981920 / 1e9:.0f} GFLOP\n"
         f"# algorithmic bytes/launch: out {WS} x 4800 B = {WS * 4800 / 1e6:.0f} MB (feature tensor), in 31.7 MB (fp16 window images, re-read from L2 per sample) + {SPL} x 87 KB weights"),
        ("tc_fc_kernel", f"prof_fc_{tag}.ncu-rep",
         f"algorithmic FLOPs/launch = {WS} x 307 456 = {WS * 307456 / 1e9:.0f} GFLOP; bytes: feature tensor read {WS * 4800 / 1e6:.0f} MB + fc weights 307 KB x 79 tiles x {SPL} samples = {0.307 * 79 * SPL:.0f} MB (L2)")):
        m, tot, stalls, mnem = capture(os.path.join(G, rep))
        f.write(f"# {rnd} -- ncu --set full --clock-control none --import-source on, {name}<false> (launch 3 of bench.py --steps 1 --warmup 3 --no-cpu --no-train)\n# {head}\n")
        for k in keys:
            if k in m:
                f.write(f"{k} [{m[k][1]}]: {m[k][0]}\n")
        f.write("warp-stall sampling (all samples): " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(stalls.items(), key=lambda x: -x[1])[:8]) + "\n")
        f.write("async-unit SASS executed (warp-instructions): " + ", ".join(f"{k} {v}" for k, v in sorted(mnem.items())) + "\n\n")
        if name == "tc_conv_kernel":
            db = bytes_of(*m['dram__bytes_read.sum']) + bytes_of(*m['dram__bytes_write.sum'])
            json.dump({"kernel": name, "launch": f"{SPL} MC samples x 10000 windows ({WS} window-samples)", "dram_bytes_per_launch": db,
                       "dram_bytes_per_window_sample": db / WS, "algorithmic_bytes_per_window_sample": 4800 + 2160 / 100,
                       "source": f"profiles/{rnd}_ncu_tc_kernels.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"},
                      open(os.path.join(P, "conv_traffic.json"), "w"), indent=1)
print(open(os.path.join(P, f"{rnd}_ncu_tc_kernels.txt")).read())
print(open(os.path.join(P, f"{rnd}_launches_predict_tc.txt")).read())

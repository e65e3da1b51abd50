#!/usr/bin/env python
"""Tracked summary of the MC-dropout conv-kernel captures (tools/profile_mcd.py under ncu --set full): before / after the round-2
epilogue changes.  usage: python tools/make_profiles_mcd.py <round-prefix>"""
import collections, csv, io, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r02"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp16.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'launch__registers_per_thread']


def ncu(rep, page):
    return list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout)))


with open(os.path.join(P, f"{rnd}_ncu_mcd_conv.txt"), "w") as f:
    f.write(f"# {rnd} -- ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 4 -c 1 python tools/profile_mcd.py\n"
            "# tc_conv_kernel<true> = the conv stack with fused Philox dropout masks (configs[1]: p = 0.241437, B = 10 000 x S = 100 per launch)\n"
            "# 'before' = 16-bit decisions via the emulated integer SIMD compare, round keys derived per block, 1/keep multiply per activation\n"
            "# 'after'  = 14-bit decisions via HSET2, round keys from the kernel parameters, 1/keep folded into the packed weights (DESIGN 4.4)\n"
            "# opcode counts = 'Instructions Executed' of the source page summed per SASS mnemonic (warp instructions of one launch; the\n"
            "#   wait-loop opcodes SYNCS / NANOSLEEP / BRA vary with the replay pass)\n")
    for tag in ("before", "after"):
        rep = os.path.join(G, f"prof_mcd_conv_{tag}.ncu-rep")
        raw = ncu(rep, "raw")
        h, u, r = raw[0], raw[1], raw[2]
        f.write(f"\n== {tag}\n")
        for k in KEYS:
            if k in h:
                f.write(f"  {k:88s} {r[h.index(k)]:>16s} {u[h.index(k)]}\n")
        src = ncu(rep, "source")
        hh = src[1]
        si, ii = hh.index("Source"), hh.index("Instructions Executed")
        cnt = collections.Counter()
        for row in src[2:]:
            try:
                n = int(row[ii])
            except (ValueError, IndexError):
                continue
            m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", row[si])
            cnt[m.group(1) if m else "?"] += n
        f.write("  opcodes (M warp-instructions): " + ", ".join(f"{k} {v / 1e6:.0f}" for k, v in cnt.most_common(18)) + "\n")
print("written")

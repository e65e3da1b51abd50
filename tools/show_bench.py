#!/usr/bin/env python
"""Short summary of a bench.py JSON line: python tools/show_bench.py gpurun_out/bench_x.json"""
import json
import sys

d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
r = d["roofline"]
print(f"{d['metric']} n_gpus={d['n_gpus']}: value {d['value'] / 1e6:.1f} M ({d['ms_per_step']:.3f} ms/step), e2e {d['e2e']['value'] / 1e6:.1f} M, "
      f"gpu_launches {d['gpu_launches']}, clocks {d['clocks'].get('sm_mhz')} MHz {d['clocks'].get('reasons')}")
print(f"  roofline {r['kernel']}: {r['achieved']:.0f} / {r['peak']:.0f} {r['unit']} = {r['frac']:.3f}, {r['ms_per_launch']:.3f} ms/launch, share {r['share_of_step']}, traffic {r['traffic']}")
for k in r.get("other_kernels", []):
    print(f"  {k['kernel']}: {k['achieved']:.0f} / {k['peak']:.0f} {k['unit']} = {k['frac']:.3f}, {k['ms_per_launch']:.3f} ms/launch")
for m, v in d.get("train", {}).items():
    print(f"  train {m}: {v['ms_per_step']:.3f} ms/step = {v['windows_per_s'] / 1e3:.0f} k windows/s [{v['gemm_backend']}] {v['ms_per_step_by_backend']} "
          f"eager {v.get('ms_per_step_eager_fused', v.get('ms_per_step_eager_simt', v.get('ms_per_step_eager')))} cpu {v.get('cpu_windows_per_s')}"
          + (f" ({v['speedup_vs_cpu_port']:.0f}x)" if v.get("speedup_vs_cpu_port") else ""))
if d.get("mcd_predict"):
    print(f"  mcd {d['mcd_predict']['window_samples_per_s'] / 1e6:.1f} M")
if d.get("flipout_predict"):
    print(f"  flipout predict (S=20) {d['flipout_predict']['window_samples_per_s'] / 1e6:.1f} M")
if d.get("deep_ensemble_full"):
    f = d["deep_ensemble_full"]
    print(f"  configs[4] at scale: {f['window_units_per_s'] / 1e6:.1f} M window-units/s, {f['seconds_per_pass']:.2f} s per pass of 1M windows x (1000 + 5)")
if d.get("radial_sweep"):
    print("  radial", {k: round(v["window_samples_per_s"] / 1e6, 1) for k, v in d["radial_sweep"].items() if k != "note"})
if d.get("deep_ensemble"):
    print(f"  deep ensemble {d['deep_ensemble']['window_members_per_s'] / 1e6:.1f} M window-members/s ({d['deep_ensemble']['ms_per_step']:.3f} ms)")
if d.get("cpu_baseline"):
    print(f"  cpu baseline {d['cpu_baseline']['value']:.0f} ({d['cpu_baseline']['cores']} cores): {d['cpu_baseline']['sample']}")
if d.get("other_nets_predict"):
    print("  other nets (S=20 predict, M window-samples/s):",
          {n: {k: round(v / 1e6, 1) for k, v in r["window_samples_per_s"].items()} for n, r in d["other_nets_predict"].items()})

#!/bin/bash
# usage (on the GPU box, repo root): tools/profile_round2.sh <tag>     -- round-2 profiling pass
# Every ncu pass runs only after the same command exited 0 without ncu.  Outputs land in gpurun_out/ (scratch);
# tools/make_profiles.py (predictive path) and tools/make_profiles_fused.py (level-fused training kernels) turn them into the
# tracked summaries under profiles/.
tag=${1:-r02}
set -x
# ---- predictive path (headline): plain run, launch list, one full capture per tcgen05 kernel
python bench.py --steps 2 --warmup 3 --no-cpu --no-train > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-train > gpurun_out/ncu_list_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 2 -c 1 -o gpurun_out/prof_conv_${tag} -f \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-train > gpurun_out/ncu_conv_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_fc_kernel -s 2 -c 1 -o gpurun_out/prof_fc_${tag} -f \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-train > gpurun_out/ncu_fc_${tag}.log 2>&1
# ---- ELBO train step on the level-fused tcgen05 back-end (B = 256): eager launches (the graph replays exactly these kernels)
for mode in lrt flipout; do
  BRL_NO_GRAPH=1 python tools/profile_train.py $mode 3 fused > gpurun_out/train_fused_${mode}_${tag}.log 2>&1 || exit 1
  BRL_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv \
    --log-file gpurun_out/launches_train_fused_${mode}_${tag}.csv python tools/profile_train.py $mode 3 fused > gpurun_out/ncu_tfl_${mode}_${tag}.log 2>&1
done
BRL_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:tt_ -s 28 -c 14 -o gpurun_out/prof_train_fused_${tag} -f \
  python tools/profile_train.py lrt 3 fused > gpurun_out/ncu_tf_${tag}.log 2>&1
ls -la gpurun_out/*${tag}*

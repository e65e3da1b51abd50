#!/bin/bash
# usage (on the GPU box, repo root): tools/profile_predict.sh <tag>
# 1. plain bench (must exit 0), 2. ncu launch list of the same command, 3. one --set full capture of each tcgen05 kernel
tag=${1:-rXX}
set -x
python bench.py --steps 2 --warmup 3 --no-cpu --no-train > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-train > gpurun_out/ncu_list_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 2 -c 1 -o gpurun_out/prof_conv_${tag} -f \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-train > gpurun_out/ncu_conv_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_fc_kernel -s 2 -c 1 -o gpurun_out/prof_fc_${tag} -f \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-train > gpurun_out/ncu_fc_${tag}.log 2>&1

#!/bin/bash
# usage (on the GPU box, repo root): tools/profile_round.sh <tag>
# Every ncu pass runs only after the same command exited 0 without ncu.  Outputs land in gpurun_out/ (scratch);
# tools/make_profiles.py + tools/make_profiles_train.py turn them into the tracked summaries under profiles/.
tag=${1:-rXX}
set -x
# ---- predictive path (headline): plain run, launch list, one full capture per tcgen05 kernel
python bench.py --steps 2 --warmup 3 --no-cpu --no-train > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-train > gpurun_out/ncu_list_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 2 -c 1 -o gpurun_out/prof_conv_${tag} -f \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-train > gpurun_out/ncu_conv_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_fc_kernel -s 2 -c 1 -o gpurun_out/prof_fc_${tag} -f \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-train > gpurun_out/ncu_fc_${tag}.log 2>&1
# ---- MC-dropout predictive path (configs[1])
python tools/profile_mcd.py 25 > gpurun_out/mcd_${tag}.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 2 -c 1 -o gpurun_out/prof_conv_mcd_${tag} -f \
  python tools/profile_mcd.py 25 > gpurun_out/ncu_mcd_${tag}.log 2>&1
# ---- ELBO train step (LRT, B = 256): eager launches (the graph replays exactly these kernels)
BRL_NO_GRAPH=1 python tools/profile_train.py lrt 2 simt > gpurun_out/train_${tag}.log 2>&1 || exit 1
BRL_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_train_${tag}.csv \
  python tools/profile_train.py lrt 2 simt > gpurun_out/ncu_train_list_${tag}.log 2>&1
BRL_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 5 -c 1 -o gpurun_out/prof_train_fwd_${tag} -f \
  python tools/profile_train.py lrt 2 simt > gpurun_out/ncu_tr1_${tag}.log 2>&1
BRL_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 17 -c 1 -o gpurun_out/prof_train_dx_${tag} -f \
  python tools/profile_train.py lrt 2 simt > gpurun_out/ncu_tr2_${tag}.log 2>&1
BRL_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:conv_dw_kernel -s 4 -c 1 -o gpurun_out/prof_train_dw_${tag} -f \
  python tools/profile_train.py lrt 2 simt > gpurun_out/ncu_tr3_${tag}.log 2>&1
ls -la gpurun_out/*${tag}*

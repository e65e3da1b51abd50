#!/bin/bash
# usage (on the GPU box): tools/ncu_conv.sh <tag>   -- one --set full capture of the conv kernel of the predict bench
tag=${1:-x}
ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 4 -c 1 -o gpurun_out/prof_conv_$tag -f \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-train > gpurun_out/ncu_$tag.log 2>&1

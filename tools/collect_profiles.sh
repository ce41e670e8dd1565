#!/bin/bash
# Gathers the ncu evidence summarised under profiles/ (run on the GPU box: gpurun -- 'bash tools/collect_profiles.sh r01').
# 1. launch lists (time, DRAM bytes, warp instructions per launch) of tools/stage_times.py, one scene at a time;
# 2. one `--set full` capture of the dominant kernel (k_composite) per scene, steady-state frame.
TAG=${1:-r01}
mkdir -p gpurun_out
for scene in random100k tiger4k; do
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
      --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_${scene}.csv \
      python tools/stage_times.py $scene > gpurun_out/${TAG}_launches_${scene}.log 2>&1
  ncu --set full --import-source on --clock-control none -k regex:k_composite -s 3 -c 1 -f \
      -o gpurun_out/${TAG}_composite_${scene} python tools/stage_times.py $scene > gpurun_out/${TAG}_composite_${scene}.log 2>&1
done
ls -la gpurun_out | grep ${TAG}_

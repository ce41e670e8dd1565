#!/bin/bash
# Gathers the ncu evidence summarised under profiles/ (run on the GPU box: gpurun -- 'bash tools/collect_profiles.sh r02').
# 1. launch lists (time, DRAM bytes, warp instructions per launch) of tools/stage_times.py, one scene at a time;
# 2. one `--set full` capture per scene of the fill + tile stage (k_tile_solid, k_tile_alpha), steady-state frame;
# 3. one `--set full` capture of the bin count pass (k_bin<1>) with the L2 atomic / reduction counters the
#    north star asks for.
TAG=${1:-r02}
mkdir -p gpurun_out
for scene in random100k tiger4k; do
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
      --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_${scene}.csv \
      python tools/stage_times.py $scene > gpurun_out/${TAG}_launches_${scene}.log 2>&1
  for k in k_tile_alpha k_tile_solid; do
    ncu --set full --import-source on --clock-control none -k regex:$k -s 3 -c 1 -f \
        -o gpurun_out/${TAG}_${k}_${scene} python tools/stage_times.py $scene > gpurun_out/${TAG}_${k}_${scene}.log 2>&1
  done
done
# (k_bin<1> and k_bin<0> share the base name: the 5th / 6th launch named k_bin are the count and the emit pass of the
#  third frame, a steady-state one)
BIN_METRICS=lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,lts__t_requests_op_atom.sum,lts__t_requests_op_red.sum,l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum,l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed
for pass in count:5 emit:6; do
  ncu --set full --import-source on --clock-control none --metrics $BIN_METRICS --kernel-id ::k_bin:${pass#*:} -f \
      -o gpurun_out/${TAG}_k_bin_${pass%:*}_random100k python tools/stage_times.py random100k > gpurun_out/${TAG}_k_bin_${pass%:*}_random100k.log 2>&1
done
ls -la gpurun_out | grep ${TAG}_

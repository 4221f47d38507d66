#!/bin/bash
# Round-2 evidence capture on one B200 (run under gpurun; outputs under gpurun_out/, summaries copied to profiles/ afterwards):
#   1. launch list of one warm step            (ncu --metrics gpu__time_duration.sum)
#   2. ncu --set full of every conv launch of one forward (both instantiations, all shapes)
#   3. ncu --set full of the kNN / SC2-PCR / map kernels
#   4. compute-sanitizer racecheck + memcheck on the persistent conv loop (tools/racecheck_conv.py)
set -x
R=gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $R/launches_r02_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $R/cap_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sparse_conv_h_kernel -s 69 -c 23 -o $R/prof_conv_r02 $B > $R/cap_conv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"knn_tc_kernel|first_order_bits|seed_consensus_kernel|seed_fitness|csr_fill|nms_bits|kernel_map_sym|kernel_map_kernel|permute_columns|sparse_conv_cin1|seed_kabsch|refine_kernel|hash_build_kernel" -s 40 -c 30 -o $R/prof_rest_r02 $B > $R/cap_rest.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:power_step -s 68 -c 2 -o $R/prof_power_r02 $B > $R/cap_power.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tools/racecheck_conv.py > $R/racecheck_conv_r02.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/racecheck_conv.py > $R/memcheck_conv_r02.log 2>&1
tail -5 $R/racecheck_conv_r02.log $R/memcheck_conv_r02.log
ls -la $R/*.ncu-rep | tail -4

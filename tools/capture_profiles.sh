#!/bin/bash
# Round-2 evidence capture on one B200 (run under gpurun).  Only text summaries leave the box (gpurun_out/ is capped at 64 MiB):
#   1. launch list of warm steps                (ncu --metrics gpu__time_duration.sum)
#   2. ncu --set full of the conv launches of one forward (both instantiations, all shapes)          [CONV=1]
#   3. ncu --set full of the kNN / SC2-PCR / map / stem kernels
#   4. compute-sanitizer racecheck on the persistent conv loop (tools/racecheck_conv.py)               [RACE=1]
set -x
R=gpurun_out
T=/tmp/eyoc_prof
mkdir -p $T
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-check-gather --no-overlap"
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $R/launches_r02_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-check-gather > $R/cap_launch.log 2>&1
if [ "$CONV" = "1" ]; then
  ncu --set full --clock-control none -k regex:sparse_conv_h_kernel -s 66 -c 22 -o $T/prof_conv $B > $R/cap_conv.log 2>&1
  python tools/ncu_summary.py $T/prof_conv.ncu-rep $R/ncu_conv_r02.txt > /dev/null
fi
ncu --set full --clock-control none -k regex:"knn_tc_kernel|first_order_bits|seed_consensus_kernel|seed_fitness|csr_fill|nms_bits|kernel_map_sym|kernel_map_kernel|permute_columns|stem_ones|power_fused|block_insert|seed_kabsch|refine_kernel|hash_build_kernel|gather_rows|tile_key_kernel" -s 50 -c 40 -o $T/prof_rest $B > $R/cap_rest.log 2>&1
python tools/ncu_summary.py $T/prof_rest.ncu-rep $R/ncu_rest_r02.txt > /dev/null
if [ "$RACE" = "1" ]; then
  timeout 420 compute-sanitizer --tool racecheck --racecheck-report all python tools/racecheck_conv.py > $R/racecheck_conv_r02.log 2>&1
  echo "racecheck exit code $?" >> $R/racecheck_conv_r02.log
fi
ls -la $T $R | tail -8
du -sh $R

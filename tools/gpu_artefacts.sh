#!/bin/sh
# Developer tool (GPU box): regenerates the round-2 ncu artefacts under gpurun_out/ in one call.
#   gpurun --timeout 2400 -- sh tools/gpu_artefacts.sh
# Reports are converted to their raw CSV page on the box and deleted (gpurun copies back at most 64 MiB).
set -x
export CPB200_NO_GRAPH=1     # ncu sees the kernels one by one
B="python bench.py --steps 2 --warmup 3 --no-sub"
csv() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.csv 2>/dev/null; rm -f gpurun_out/$1.ncu-rep; }
# pile1m: settle 20 + warm-up 3 = 23 steps before the timed ones.  Kernels matching the regex per step: k_sort_scatter x3,
# k_bvh_refit, k_bvh_pairs, k_collide x3, k_colour_solve x2 (colour + rows, then iterate) = 10: one whole step is captured.
timeout 900 ncu --set full --clock-control none -k "regex:k_colour_solve|k_collide|k_bvh_refit|k_bvh_pairs|k_sort_scatter" -s 230 -c 10 -f -o gpurun_out/r02_ncu_full_step_pile1m $B --workload pile1m > gpurun_out/ncu_a.log 2>&1
csv r02_ncu_full_step_pile1m
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 200 --csv --log-file gpurun_out/r02_launches_pile1m.csv $B --workload pile1m > gpurun_out/ncu_b.log 2>&1
# GJK class on config 3 (settle 120 + 3 steps, 3 collide kernels each), space-local solver on config 5 (settle 300 + 3, 2 each)
timeout 600 ncu --set full --clock-control none -k regex:k_collide -s 371 -c 1 -f -o gpurun_out/r02_ncu_full_collide2_mixed100k $B --workload mixed100k > gpurun_out/ncu_c.log 2>&1
csv r02_ncu_full_collide2_mixed100k
timeout 600 ncu --set full --clock-control none -k regex:k_sl_solve -s 606 -c 2 -f -o gpurun_out/r02_ncu_full_sl_solve_batch $B --workload batch > gpurun_out/ncu_d.log 2>&1
csv r02_ncu_full_sl_solve_batch
ls -la gpurun_out/

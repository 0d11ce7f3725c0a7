#!/bin/sh
# Developer tool (GPU box): regenerates the round-2 measurement artefacts under gpurun_out/ in one call.
#   gpurun --timeout 2400 -- sh tools/gpu_artefacts.sh
set -x
export CPB200_NO_GRAPH=1     # ncu sees the kernels one by one
B="python bench.py --steps 2 --warmup 3 --no-sub"
# pile1m: settle 20 + warm-up 3 = 23 steps before the timed ones
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 200 --csv --log-file gpurun_out/r02_launches_pile1m.csv $B --workload pile1m > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_colour_solve -s 47 -c 1 -f -o gpurun_out/r02_ncu_iterate $B --workload pile1m > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_colour_solve -s 46 -c 1 -f -o gpurun_out/r02_ncu_colour_rows $B --workload pile1m > gpurun_out/ncu_c.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_collide -s 69 -c 1 -f -o gpurun_out/r02_ncu_collide0 $B --workload pile1m > gpurun_out/ncu_d.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_bvh_refit -s 23 -c 1 -f -o gpurun_out/r02_ncu_bvh_refit $B --workload pile1m > gpurun_out/ncu_e.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_bvh_pairs -s 23 -c 1 -f -o gpurun_out/r02_ncu_bvh_pairs $B --workload pile1m > gpurun_out/ncu_f.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_sort_scatter -s 69 -c 1 -f -o gpurun_out/r02_ncu_sort_scatter $B --workload pile1m > gpurun_out/ncu_g.log 2>&1
# GJK class on config 3 (settle 120 + 3), space-local solver on config 5 (settle 300 + 3)
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_collide -s 371 -c 1 -f -o gpurun_out/r02_ncu_collide2_mixed100k $B --workload mixed100k > gpurun_out/ncu_h.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_sl_solve -s 606 -c 1 -f -o gpurun_out/r02_ncu_sl_solve_batch $B --workload batch > gpurun_out/ncu_i.log 2>&1
ls -la gpurun_out/*.ncu-rep

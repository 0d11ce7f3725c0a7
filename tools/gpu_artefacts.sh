set -x
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r01_bench_pile1m_reference.json 2> gpurun_out/bench_ref.err
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r01_bench_pile1m.json 2> gpurun_out/bench.err
timeout 600 python bench.py --steps 30 --warmup 5 --workload batch > gpurun_out/r01_bench_batch4096_n1.json 2> gpurun_out/bench_batch.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_pile1m.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_b.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_colour_solve -s 48 -c 1 -f -o gpurun_out/r01_ncu_solve python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_c.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_collide -s 144 -c 1 -f -o gpurun_out/r01_ncu_collide python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r01_*
cat gpurun_out/r01_bench_pile1m.json | cut -c1-600

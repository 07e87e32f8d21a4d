#!/bin/bash
# round-2 evidence run (1 GPU): bench lines, reference arm, ncu launch list and full captures
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; echo rc=$?
python bench.py --workload C3 --steps 20 --warmup 5 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; echo rc=$?
python bench.py --workload C4s8 --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/r2_bench_c4s8.json 2>/dev/null; echo rc=$?
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_c4_reference.json 2> gpurun_out/r2_ref.err; echo rc=$?
python bench.py --workload C5 --steps 5 --warmup 2 > gpurun_out/r2_bench_c5.json 2>/dev/null; echo rc=$?
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity"
$B > /dev/null 2>&1; echo plain rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c4.csv $B > gpurun_out/ncu_l.log 2>&1; echo rc=$?
ncu --set full --clock-control none -k regex:"k_stats_units|k_stats_incr|k_table" --launch-skip 20 --launch-count 6 -o gpurun_out/r2_stats_table_c4 $B > gpurun_out/ncu_s.log 2>&1; echo rc=$?
B3="python bench.py --workload C3 --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity"
$B3 > /dev/null 2>&1; echo plain rc=$?
ncu --set full --clock-control none -k regex:"k_cells_sell|k_stats_stream|k_stats_incr|k_table" --launch-skip 24 --launch-count 8 -o gpurun_out/r2_all_c3 $B3 > gpurun_out/ncu_c3.log 2>&1; echo rc=$?
python -m pytest tests -m gpu -q > gpurun_out/pytest_r2_final.log 2>&1; tail -3 gpurun_out/pytest_r2_final.log

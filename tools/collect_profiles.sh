#!/bin/bash
# Round-2 evidence run on ONE B200 (about 7 minutes): the bench lines behind profiles/README.md, the reference arm, the
# ncu launch list of the bench command and `ncu --set full` captures of every sweep kernel at C4 and C3.  Every ncu
# command is run only after the same command has exited 0 without ncu.  Outputs go to gpurun_out/ (scratch); the
# summaries that are tracked are made from them by tools/summarize_profiles.py.
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; echo rc=$?
python bench.py --workload C3 --steps 20 --warmup 5 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; echo rc=$?
python bench.py --workload C4s8 --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/r2_bench_c4s8.json 2>/dev/null; echo rc=$?
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_c4_reference.json 2> gpurun_out/r2_ref.err; echo rc=$?
python bench.py --impl reference --workload C3 --steps 20 --warmup 5 > gpurun_out/r2_bench_c3_reference.json 2>> gpurun_out/r2_ref.err; echo rc=$?
python bench.py --workload C5 --steps 5 --warmup 2 > gpurun_out/r2_bench_c5.json 2>/dev/null; echo rc=$?
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity"
$B > /dev/null 2>&1; echo plain rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c4.csv $B > gpurun_out/ncu_l.log 2>&1; echo rc=$?
# the first statistic launch of a run is a full pass (forced); later ones are the incremental correction
ncu --set full --clock-control none -k regex:k_stats_units --launch-count 1 -o gpurun_out/r2_stats_full_c4 -f $B > gpurun_out/ncu_s1.log 2>&1; echo rc=$?
ncu --set full --clock-control none -k regex:"k_cells_sell|k_stats_units|k_table" --launch-skip 12 --launch-count 3 -o gpurun_out/r2_sweep_c4 -f $B > gpurun_out/ncu_s2.log 2>&1; echo rc=$?
B3="python bench.py --workload C3 --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity"
$B3 > /dev/null 2>&1; echo plain rc=$?
ncu --set full --clock-control none -k regex:k_stats_stream --launch-count 1 -o gpurun_out/r2_stats_full_c3 -f $B3 > gpurun_out/ncu_c31.log 2>&1; echo rc=$?
ncu --set full --clock-control none -k regex:"k_cells_sell|k_stats_stream|k_table" --launch-skip 12 --launch-count 3 -o gpurun_out/r2_sweep_c3 -f $B3 > gpurun_out/ncu_c32.log 2>&1; echo rc=$?
python tools/e2e_breakdown.py > gpurun_out/r2_e2e_breakdown.txt 2>&1; echo rc=$?
# device timeline of the last two sweeps (%globaltimer stamps inside the kernels) and the duration of every sweep
for w in C4 C4s8 C3; do CDL_TIMELINE=1 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu --no-e2e --no-parity 2>&1 >/dev/null | grep "cdl " | head -45 > gpurun_out/r2_timeline_$w.txt; done
# gpurun_out/ comes back only if it stays below 64 MiB: summarise the reports ON the box (ncu reads them without a GPU too),
# bring the summaries back, drop the reports
python tools/summarize_profiles.py > gpurun_out/summarize.log 2>&1; echo rc=$?
cp profiles/r2_ncu_full_c4_summary.txt profiles/r2_ncu_full_c3_summary.txt profiles/ncu_traffic.json gpurun_out/
rm -f gpurun_out/*.ncu-rep
if [ "$1" = "--tests" ]; then python -m pytest tests -m gpu -q > gpurun_out/pytest_r2_final.log 2>&1; tail -3 gpurun_out/pytest_r2_final.log; fi
du -sh gpurun_out

#!/bin/bash
# Round-2 measurement run on the GPU box (gpurun): bench lines, ncu launch list, ncu --set full of the
# top kernels, timeline of the scheduler's GPU idle time, sanitizer passes.  Outputs under gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/r2_clocks.csv &
SMI=$!
python bench.py > $O/r2_bench_C3.json 2> $O/r2_bench_C3.err
python bench.py --steps 20 --warmup 5 > $O/r2_bench_C3_driver_cmd.json 2>> $O/r2_bench_C3.err
BNPC_LOCKSTEP=1 python bench.py --steps 20 --warmup 5 --group-size 8 --no-extras --no-cpu-baseline > $O/r2_bench_C3_lockstep8.json 2>> $O/r2_bench_C3.err
kill $SMI
ncu --metrics gpu__time_duration.sum --clock-control none --print-kernel-base demangled -c 6000 --csv --log-file $O/r2_launches_C3.csv \
    python bench.py --steps 10 --warmup 3 --windows 1 --no-extras --no-cpu-baseline > $O/r2_launches_C3.log 2>&1
BNPC_LOCKSTEP=1 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'suffstat|gibbs_exact_kernel|gibbs_sweep|ll_matrix_i8' -s 40 -c 12 -o $O/r2_top \
    python bench.py --steps 4 --warmup 3 --windows 1 --no-extras --no-cpu-baseline --group-size 8 > $O/r2_top.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_group.py -x -q -k "independent or lockstep" > $O/r2_memcheck_group.log 2>&1
timeout 600 compute-sanitizer --tool racecheck --racecheck-report hazard --print-limit 20 --kernel-regex kns=gibbs_sweep python tools/race_repro.py 15 11 1 > $O/r2_racecheck_sweep_after_fix.log 2>&1
tail -3 $O/r2_memcheck_group.log; tail -2 $O/r2_racecheck_sweep_after_fix.log
ls -la $O | tail -20

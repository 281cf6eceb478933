#!/bin/bash
# Final measurement run of round 2 on the GPU box (gpurun): GPU tests, the driver's bench command,
# the ncu launch list of the bench command, ncu --set full of the shared integer tcgen05 rows and
# the other top kernels.  Outputs under gpurun_out/ (summaries are made from them with
# tools/profile_summary.py and committed under profiles/).  `full` as first argument adds the
# default bench run (100 timed steps per window).
set -u
O=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > $O/r2f_tests.log 2>&1; tail -3 $O/r2f_tests.log
python bench.py --steps 20 --warmup 5 > $O/r2f_bench_C3_driver_cmd.json 2> $O/r2f_bench_C3.err
ncu --metrics gpu__time_duration.sum --clock-control none --print-kernel-base demangled -c 6000 --csv --log-file $O/r2f_launches_C3.csv \
    python bench.py --steps 10 --warmup 3 --windows 1 --no-extras --no-cpu-baseline > $O/r2f_launches_C3.log 2>&1
BNPC_LOCKSTEP=1 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'ll_matrix_i8s|gibbs_exact_kernel|mh_theta_kernel|suffstat_kernel|ll_few_kernel' -s 30 -c 14 -o $O/r2f_top \
    python bench.py --steps 6 --warmup 3 --windows 1 --no-extras --no-cpu-baseline --group-size 8 > $O/r2f_top.log 2>&1
if [ "${1:-}" = "full" ]; then python bench.py > $O/r2f_bench_C3.json 2>> $O/r2f_bench_C3.err; fi
tail -c 300 $O/r2f_bench_C3_driver_cmd.json

#!/bin/bash
set -u
O=gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -x -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > $O/r2_bench_C3_driver_cmd.json 2> $O/r2_bench_C3.err
python bench.py > $O/r2_bench_C3.json 2>> $O/r2_bench_C3.err
ncu --metrics gpu__time_duration.sum --clock-control none --print-kernel-base demangled -c 6000 --csv --log-file $O/r2_launches_C3.csv \
    python bench.py --steps 10 --warmup 3 --windows 1 --no-extras --no-cpu-baseline > $O/r2_launches_C3.log 2>&1
BNPC_LOCKSTEP=1 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'gibbs_exact_kernel|gibbs_sweep|ll_matrix_i8|mh_theta' -s 400 -c 8 -o $O/r2_top2 \
    python bench.py --steps 6 --warmup 3 --windows 1 --no-extras --no-cpu-baseline --group-size 8 > $O/r2_top2.log 2>&1
time python bench.py --impl reference --steps 20 --warmup 5 > $O/r2_bench_C3_reference.json 2> $O/r2_bench_C3_reference.err
tail -c 1500 $O/r2_bench_C3_reference.json

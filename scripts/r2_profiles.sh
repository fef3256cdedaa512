#!/bin/bash
# round-2 profile set (run under gpurun): bench line, ncu launch list of the same command, full captures of K3 and K1
mkdir -p gpurun_out/r2p
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-continuation > gpurun_out/r2p/bench_short.json 2> gpurun_out/r2p/bench_short.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-continuation > gpurun_out/r2p/ncu_launch.log 2>&1
CATINT_PHASES=1 python scripts/profile_case.py 1000000 1024 2>&1 | grep -v transport.info > gpurun_out/r2p/phase_cycles.txt
ncu --set full --import-source on --clock-control none -k regex:pnp_bdf -c 1 -o gpurun_out/r2p/bdf_full -f \
    python scripts/profile_case.py 1000000 1024 > gpurun_out/r2p/ncu_bdf.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:pnp_rhs_kernel -s 3 -c 1 -o gpurun_out/r2p/rhs_full -f \
    python scripts/time_rhs.py > gpurun_out/r2p/ncu_rhs.log 2>&1
ls -la gpurun_out/r2p

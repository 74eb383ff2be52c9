#!/bin/bash
# One gpurun call that regenerates the evidence under gpurun_out/<tag>_*: GPU tests, smoke, both bench arms, the ncu
# launch list of the bench command and `ncu --set full` captures per launch shape (single stream / 64 streams), plus
# the warm-cache DRAM traffic of the running chain.
#   gpurun --timeout 2400 -- 'bash tools/gpu_evidence.sh r02z'
# Every ncu pass runs only after the same command has exited 0 without ncu.
tag=${1:-r02x}
what=${2:-all}
out=gpurun_out
mkdir -p $out
set -x
if [[ $what == all || $what == tests ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q -s > $out/${tag}_gpu_tests.log 2>&1; echo "tests rc=$?" >> $out/${tag}_gpu_tests.log
  tail -3 $out/${tag}_gpu_tests.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.log
fi
if [[ $what == all || $what == bench ]]; then
  timeout 900 python bench.py --impl reference > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
  timeout 1200 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
  tail -c 600 $out/${tag}_bench.err
  timeout 600 python bench.py --steps 20 --warmup 5 --no-sublegs > $out/${tag}_bench_driver_args.json 2>> $out/${tag}_bench.err
  timeout 900 python bench.py --workload c5 > $out/${tag}_bench_c5.json 2>> $out/${tag}_bench.err
  timeout 600 python tools/e2e_breakdown.py > $out/${tag}_e2e_breakdown.txt 2>&1
fi
if [[ $what == all || $what == ncu ]]; then
  # launch list of the bench command (single-stream legs only)
  timeout 600 python bench.py --steps 30 --warmup 3 --streams 0 --no-cpu --no-sublegs > $out/${tag}_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $out/${tag}_launches_c2_single.csv \
      python bench.py --steps 30 --warmup 3 --streams 0 --no-cpu --no-sublegs > $out/${tag}_ncu_launches.log 2>&1
  # steady-state frames launch 8 kernels (single stream, speculative matching on) or 9 (64 streams): a window of two
  # frames' worth of consecutive launches holds every kernel of the chain exactly twice
  for S in 1 64; do
    n=$([[ $S == 1 ]] && echo 16 || echo 18)
    timeout 300 python tools/profile_target.py --streams $S --frames 8 > $out/${tag}_pt$S.log 2>&1 && \
    timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 30 -c $n \
        -o $out/${tag}_s$S -f python tools/profile_target.py --streams $S --frames 8 > $out/${tag}_ncu_s$S.log 2>&1
    ncu -i $out/${tag}_s$S.ncu-rep --page raw --csv > $out/${tag}_s$S.raw.csv 2>/dev/null
  done
  python tools/ncu_summary.py traffic $out/${tag}_s1.raw.csv > $out/${tag}_traffic_cold_c2.json
  # warm-cache DRAM traffic of the chain: counters only, caches left alone
  timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum \
      --launch-skip 1400 -c 32 --csv --log-file $out/${tag}_warm_s1.csv python tools/profile_target.py --streams 1 --frames 200 \
      > $out/${tag}_ncu_warm_s1.log 2>&1
  timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum \
      --launch-skip 27 -c 36 --csv --log-file $out/${tag}_warm_s64.csv python tools/profile_target.py --streams 64 --frames 8 \
      > $out/${tag}_ncu_warm_s64.log 2>&1
  python tools/ncu_summary.py traffic_warm $out/${tag}_warm_s1.csv $out/${tag}_traffic_cold_c2.json > $out/${tag}_roofline_traffic_c2.json
  python tools/ncu_summary.py traffic_warm $out/${tag}_warm_s64.csv > $out/${tag}_traffic_warm_s64.json
fi
ls -la $out | tail -30

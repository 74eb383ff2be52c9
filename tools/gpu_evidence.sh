#!/bin/bash
# One gpurun call that regenerates the evidence under gpurun_out/<tag>_*: GPU tests, both bench arms, the ncu
# launch list of the bench command and one `ncu --set full` capture per launch shape (single stream / 64 streams).
#   gpurun --timeout 1500 -- 'bash tools/gpu_evidence.sh r01b'
# Every ncu pass runs only after the same command has exited 0 without ncu.
tag=${1:-r01x}
what=${2:-all}
out=gpurun_out
mkdir -p $out
set -x
if [[ $what == all || $what == tests ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1; echo "tests rc=$?" >> $out/${tag}_tests.log
  tail -3 $out/${tag}_tests.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.log
fi
if [[ $what == all || $what == bench ]]; then
  timeout 600 python bench.py --impl reference --steps 100 --warmup 10 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
  timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
  tail -c 600 $out/${tag}_bench.err
fi
if [[ $what == all || $what == ncu ]]; then
  # launch list of the bench command (single stream leg only, no CPU leg)
  timeout 600 python bench.py --steps 30 --warmup 3 --streams 0 --no-cpu > $out/${tag}_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/${tag}_launches.csv \
      python bench.py --steps 30 --warmup 3 --streams 0 --no-cpu > $out/${tag}_ncu_launches.log 2>&1
  for S in 1 64; do
    timeout 300 python tools/profile_target.py --streams $S --frames 5 > $out/${tag}_pt$S.log 2>&1 && \
    timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 24 -c 16 \
        -o $out/${tag}_s$S -f python tools/profile_target.py --streams $S --frames 5 > $out/${tag}_ncu_s$S.log 2>&1
    ncu -i $out/${tag}_s$S.ncu-rep --page raw --csv > $out/${tag}_s$S.raw.csv 2>/dev/null
  done
  python tools/ncu_summary.py traffic $out/${tag}_s1.raw.csv > $out/${tag}_traffic_cold_c2.json
  # warm-cache DRAM traffic of the chain: counters only, caches left alone
  # S=1: 200 distinct frames (144 MB > L2, like bench.py), counters taken from frame 191 on; S=64: 7 steps of 46 MB
  timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum \
      --launch-skip 1527 -c 32 --csv --log-file $out/${tag}_warm_s1.csv python tools/profile_target.py --streams 1 --frames 200 \
      > $out/${tag}_ncu_warm_s1.log 2>&1
  timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum \
      --launch-skip 24 -c 40 --csv --log-file $out/${tag}_warm_s64.csv python tools/profile_target.py --streams 64 --frames 7 \
      > $out/${tag}_ncu_warm_s64.log 2>&1
  python tools/ncu_summary.py traffic_warm $out/${tag}_warm_s1.csv $out/${tag}_traffic_cold_c2.json > $out/${tag}_roofline_traffic_c2.json
  python tools/ncu_summary.py traffic_warm $out/${tag}_warm_s64.csv > $out/${tag}_traffic_warm_s64.json
fi
ls -la $out | tail -20

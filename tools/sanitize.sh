#!/bin/bash
# compute-sanitizer passes over the frame chain (one gpurun call): memcheck (out-of-bounds / misaligned accesses, incl. the
# TMA boxes), racecheck (shared-memory hazards inside a CTA), synccheck (barrier misuse) and initcheck (reads of device
# memory nobody wrote), each over a single-stream run (latency-mode kernels, speculative matching, split graphs) and an
# 8-stream run (throughput-mode kernels).  Logs: gpurun_out/<tag>_sanitize_<tool>_s<S>.log
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh r02'
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
for S in 1 8; do
  for tool in memcheck racecheck synccheck initcheck; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 \
        python tools/profile_target.py --streams $S --frames 4 > $out/${tag}_sanitize_${tool}_s$S.log 2>&1
    echo "$tool S=$S rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/${tag}_sanitize_${tool}_s$S.log | tail -1)"
  done
done
# the two-point RANSAC kernel (config C3 at reduced size is enough: same code path)
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 \
      python tools/profile_target.py --streams 1 --frames 3 --workload c3 > $out/${tag}_sanitize_${tool}_c3.log 2>&1
  echo "$tool c3 rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/${tag}_sanitize_${tool}_c3.log | tail -1)"
done

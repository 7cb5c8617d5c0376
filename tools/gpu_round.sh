#!/bin/bash
# One GPU-box visit: test suite, bench line, kernel timings, ncu launch list and full captures.
# usage: tools/gpu_round.sh <tag> [parts for collect_profiles.sh]
tag=${1:-r02}; shift
out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt
bash tools/gpu_check.sh $tag
(timeout 900 python bench.py > $out/bench_line.json 2> $out/bench_err.log; echo "bench rc=$?" >> $out/bench_err.log); tail -c 3000 $out/bench_line.json; tail -3 $out/bench_err.log
bash tools/collect_profiles.sh ${@:-launches lu assemble nbody getrs} > $out/collect.log 2>&1; tail -20 $out/collect.log

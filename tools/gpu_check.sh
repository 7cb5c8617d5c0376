#!/bin/bash
# usage: tools/gpu_check.sh <tag>  -- GPU test suite + kernel timings + config profile into gpurun_out/<tag>/
tag=${1:-check}; out=gpurun_out/$tag; mkdir -p $out
(timeout 700 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log); tail -8 $out/pytest.log
for r in 1 8 16 64; do timeout 200 python tools/run_stage.py --stage getrs --n 20164 --nrhs $r 2>&1 | grep "getrs nrhs" | tail -2 | head -1; done | tee $out/getrs20k.txt
timeout 200 python tools/run_stage.py --stage getrs --n 4000 --nrhs 1 2>&1 | grep "getrs nrhs" | tail -2 | head -1 | tee $out/getrs4k.txt
timeout 300 python tools/prof_configs.py 2>&1 | tail -1 | tee $out/prof1.json

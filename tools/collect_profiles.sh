#!/bin/bash
# Collects the ncu / compute-sanitizer evidence of round 2 on the GPU box into gpurun_out/prof_r02/.
# usage: tools/collect_profiles.sh [part ...]   parts: launches lu assemble nbody getrs sanitize
out=gpurun_out/prof_r02; mkdir -p $out
parts=${@:-launches lu assemble nbody getrs sanitize}
NCU="ncu --clock-control none"
raw() { ncu -i $1 --page raw --csv > ${1%.ncu-rep}.raw.csv 2>/dev/null; ncu -i $1 --page details --csv > ${1%.ncu-rep}.details.csv 2>/dev/null; }
for p in $parts; do case $p in
launches)
  $NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file $out/bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sharded > $out/bench_under_ncu.log 2>&1 ;;
lu)
  $NCU --set full --kernel-name-base mangled -k regex:update_kernel_tILb1 -c 2 -f -o $out/lu_update_tri python tools/run_stage.py --stage getrf --n 20164 --sym 1 > $out/lu_update_tri.log 2>&1; raw $out/lu_update_tri.ncu-rep
  $NCU --set full -k regex:diag_kernel_symb -s 20 -c 1 -f -o $out/lu_diag python tools/run_stage.py --stage getrf --n 5300 --sym 1 > $out/lu_diag.log 2>&1; raw $out/lu_diag.ncu-rep
  $NCU --set full -k regex:update_lat_kernel -s 2 -c 1 -f -o $out/lu_update_lat python tools/run_stage.py --stage getrf --n 5300 --sym 1 > $out/lu_update_lat.log 2>&1; raw $out/lu_update_lat.ncu-rep
  $NCU --set full -k regex:trsm_sym_lat_kernel -s 2 -c 1 -f -o $out/lu_trsm_lat python tools/run_stage.py --stage getrf --n 5300 --sym 1 > $out/lu_trsm_lat.log 2>&1; raw $out/lu_trsm_lat.ncu-rep ;;
assemble)
  $NCU --set full -k regex:assemble_dense -c 1 -f -o $out/assemble_dense python tools/run_stage.py --stage getrf --n 20164 --sym 1 > $out/assemble.log 2>&1; raw $out/assemble_dense.ncu-rep ;;
nbody)
  $NCU --set full -k regex:nbody_kernel -c 4 -f -o $out/nbody python tools/run_stage.py --stage nbody --n 20164 > $out/nbody.log 2>&1; raw $out/nbody.ncu-rep ;;
getrs)
  $NCU --set full -k regex:trsv_sweep -c 2 -f -o $out/getrs1 python tools/run_stage.py --stage getrs --n 20164 --nrhs 1 > $out/getrs1.log 2>&1; raw $out/getrs1.ncu-rep
  $NCU --set full -k regex:trsm_sweep -c 2 -f -o $out/getrs64 python tools/run_stage.py --stage getrs --n 20164 --nrhs 64 > $out/getrs64.log 2>&1; raw $out/getrs64.ncu-rep ;;
sanitize)
  for tool in memcheck racecheck synccheck; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_case.py > $out/sanitizer_$tool.log 2>&1; echo "$tool rc=$?" >> $out/sanitizer_$tool.log; tail -4 $out/sanitizer_$tool.log
  done ;;
esac; done
rm -f $out/*.ncu-rep
python tools/ncu_summary.py $out/lu_diag.raw.csv $out/lu_update_lat.raw.csv $out/lu_trsm_lat.raw.csv > $out/lat_kernels_ncu.txt 2>/dev/null
ls -la $out

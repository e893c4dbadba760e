#!/bin/bash
# compute-sanitizer over tools/sanitizer_target.py with the three tools; writes gpurun_out/r2_sanitizer.txt
out=gpurun_out/r2_sanitizer.txt
echo "compute-sanitizer on tools/sanitizer_target.py (B200, final build): every kernel variant incl. k_step_flexr (0 / 1 / 2 filter stages, hold transitions through both gap fits, independent robots) and k_step_flex (CDPR_FLEX_CLASSIC, leg model), cdpr_update, both kinematics kernels, rollouts, checkpoints, ragged batches" > $out
for tool in memcheck racecheck initcheck; do
  echo "== $tool" >> $out
  timeout 1500 compute-sanitizer --tool $tool python tools/sanitizer_target.py > gpurun_out/r2_san_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r2_san_$tool.log >> $out
  grep "^ok" gpurun_out/r2_san_$tool.log | tr '\n' ';' >> $out; echo >> $out
done
cat $out

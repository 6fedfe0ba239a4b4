# A/B timing of library variants under build/ (SDDC_B200_LIB) [+ one ncu capture of the shipped library: NCU_MODE=step|jvp|diag]
mkdir -p gpurun_out
timeout 180 python tools/stage_times.py > gpurun_out/stages_default.log 2>&1
for v in build/libsddc_*.so; do
  [ -f "$v" ] && SDDC_B200_LIB=$PWD/$v timeout 180 python tools/stage_times.py > gpurun_out/stages_$(basename $v .so).log 2>&1
done
grep -H "^step\|^jvp cached" gpurun_out/stages_*.log
if [ -n "$NCU_MODE" ]; then
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/prof_exp python tools/profile_step.py $NCU_MODE > gpurun_out/prof_exp.log 2>&1
tail -n 3 gpurun_out/prof_exp.log
fi

mkdir -p gpurun_out
timeout 120 python tools/gpu_check.py cfg3_member small_sym > gpurun_out/check_fft.log 2>&1; echo "check rc=$?"
timeout 120 python tools/stage_times.py > gpurun_out/stages_fft.log 2>&1; echo "stages rc=$?"
cat gpurun_out/check_fft.log | cut -c1-400; head -n 1 gpurun_out/stages_fft.log; tail -n 1 gpurun_out/stages_fft.log

mkdir -p gpurun_out
timeout 120 python tools/stage_times.py > gpurun_out/stages_fft.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
head -n 1 gpurun_out/stages_fft.log; tail -n 1 gpurun_out/stages_fft.log; tail -n 1 gpurun_out/pytest_gpu.log

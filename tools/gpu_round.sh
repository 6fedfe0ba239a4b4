mkdir -p gpurun_out
python tools/stage_times.py > gpurun_out/stages_fft.log 2>&1
tail -n 2 gpurun_out/stages_fft.log

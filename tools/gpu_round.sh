# one GPU-box round trip: parity report, stage timings, GPU test suite, racecheck of the FFT kernels
mkdir -p gpurun_out
python tools/gpu_check.py cfg3_member > gpurun_out/check_fft.log 2>&1
python tools/stage_times.py > gpurun_out/stages_fft.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
SANITIZE_FFT=2 timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/san_race.log 2>&1
SANITIZE_FFT=1 timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/san_mem.log 2>&1
for f in check_fft stages_fft pytest_gpu san_race san_mem; do echo "== $f"; tail -n 4 gpurun_out/$f.log; done

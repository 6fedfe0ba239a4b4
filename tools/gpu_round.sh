# one GPU-box round trip: stage timings, the whole GPU test suite
mkdir -p gpurun_out
python tools/stage_times.py > gpurun_out/stages_fft.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
for f in stages_fft pytest_gpu; do echo "== $f"; tail -n 6 gpurun_out/$f.log; done

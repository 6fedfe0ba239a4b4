set -x
mkdir -p gpurun_out
python tools/gpu_check.py cfg3_member > gpurun_out/check_fft.log 2>&1
python tools/stage_times.py > gpurun_out/stages_fft.log 2>&1
SDDC_FFT=0 python tools/stage_times.py > gpurun_out/stages_dense.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_parity.log 2>&1
tail -5 gpurun_out/check_fft.log gpurun_out/stages_fft.log gpurun_out/stages_dense.log gpurun_out/pytest_parity.log

# one GPU-box round trip: parity report, stage timings (FFT and dense transforms), parity tests, sanitizer, ncu capture
mkdir -p gpurun_out
python tools/gpu_check.py cfg3_member > gpurun_out/check_fft.log 2>&1
python tools/stage_times.py > gpurun_out/stages_fft.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_parity.log 2>&1
SANITIZE_FFT=1 timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/san_race.log 2>&1
SANITIZE_FFT=1 timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/san_mem.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/prof_fft python tools/profile_step.py > gpurun_out/prof_fft.log 2>&1
for f in check_fft stages_fft pytest_parity san_race san_mem; do echo "== $f"; tail -n 4 gpurun_out/$f.log; done

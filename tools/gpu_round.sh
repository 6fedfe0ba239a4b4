# one GPU-box round trip: GPU test suite, racecheck of the FFT kernels, bench at the config-5 shape
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
SANITIZE_FFT=2 timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/san_race.log 2>&1
python bench.py --steps 50 --warmup 5 --N_r 40 --N_fm 512 --members-per-gpu 512 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
tail -n 2 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/san_race.log
python - <<'PY'
import json
for f in ("bench_cfg5",):
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "jvp", d["jvp"]["value"], d["jvp"].get("uncached"), "diag", d["with_diagnostics"]["value"], d["stage_ms"])
PY

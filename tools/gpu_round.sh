# one GPU-box round trip: GPU test suite, bench at the config-5 shape
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
python bench.py --steps 50 --warmup 5 --N_r 40 --N_fm 512 --members-per-gpu 512 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
tail -n 3 gpurun_out/pytest_gpu.log
python - <<'PY'
import json
for f in ("bench_cfg5",):
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "jvp", d["jvp"]["value"], d["jvp"].get("uncached"), "diag", d["with_diagnostics"]["value"])
PY

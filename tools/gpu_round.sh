mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_1gpu_nocpu.json 2> gpurun_out/bench_1gpu_nocpu.err
tail -n 2 gpurun_out/pytest_gpu.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_1gpu_nocpu.json'))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "diag", d["with_diagnostics"])
PY

# one GPU-box round trip: parity tests, bench (headline shape)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_parity.log 2>&1
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
for f in pytest_parity; do echo "== $f"; tail -n 12 gpurun_out/$f.log; done
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_1gpu.json'))
for k in ("value","ms_per_step","e2e","gpu_launches","stage_ms","jvp","with_diagnostics","clocks"):
    print(k, json.dumps(d.get(k))[:400])
PY

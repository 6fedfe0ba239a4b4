# Short GPU-box round trip while a kernel is being changed: smoke, parity tests, stage times, a short bench (every piece
# under its own timeout so that a hung kernel cannot hold the box).
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -x -q > gpurun_out/pytest_parity.log 2>&1; echo "parity rc=$?"
timeout 180 python tools/stage_times.py > gpurun_out/stages.log 2>&1; echo "stages rc=$?"
timeout 400 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
tail -n 1 gpurun_out/smoke.log; tail -n 15 gpurun_out/pytest_parity.log
tail -n 7 gpurun_out/stages.log
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_quick.json'))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "jvp", d["jvp"]["value"], d["jvp"].get("uncached"),
      "diag", d["with_diagnostics"]["value"], d["clocks"], d["gpu_launches"], d["roofline"]["frac"], d.get("parity_check"))
print(d["stage_ms"])
PY

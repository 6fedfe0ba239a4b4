# One GPU-box round trip (gpurun -- 'bash tools/gpu_round.sh'): smoke, the GPU test suite, the headline bench.
# Results land in gpurun_out/ (scratch); copy what should be kept into profiles/.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
python tools/stage_times.py > gpurun_out/stages.log 2>&1
for v in build/libsddc_*.so; do
  [ -f "$v" ] && SDDC_B200_LIB=$PWD/$v python tools/stage_times.py > gpurun_out/stages_$(basename $v .so).log 2>&1
done
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
tail -n 1 gpurun_out/smoke.log; tail -n 3 gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/stages*.log
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_1gpu.json'))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "jvp", d["jvp"]["value"], d["jvp"].get("uncached"),
      "diag", d["with_diagnostics"]["value"], d["clocks"], d["gpu_launches"], d["roofline"]["frac"], d.get("cpu_baseline", {}).get("value"))
print(d["stage_ms"])
PY

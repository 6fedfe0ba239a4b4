# Bench lines on one GPU: headline (with CPU baseline), reference arm, config-2 / config-5 shapes, configs 4-5 at size.
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; echo "ref rc=$?"
python bench.py --steps 200 --warmup 5 --N_r 20 --N_fm 128 --members-per-gpu 1024 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "cfg2 rc=$?"
python bench.py --steps 50 --warmup 5 --N_r 40 --N_fm 512 --members-per-gpu 512 --no-cpu-baseline --strong-members 0 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "cfg5 rc=$?"
timeout 900 python tools/run_configs45.py > gpurun_out/configs45.log 2>&1; echo "cfg45 rc=$?"
python - <<'PY'
import json
for f in ("bench_1gpu", "bench_cfg2", "bench_cfg5"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "jvp", d["jvp"]["value"], d["jvp"].get("uncached"),
              "diag", d["with_diagnostics"]["value"], d["clocks"], d["gpu_launches"], d["roofline"]["frac"], (d.get("parity_check") or {}).get("ok"),
              (d.get("strong_scaling") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
print(open("gpurun_out/bench_reference_arm.json").read()[:600])
PY
tail -c 1500 gpurun_out/configs45.log

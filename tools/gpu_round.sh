# one GPU-box round trip: smoke, GPU test suite, sanitizer, benches, launch list, ncu captures
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
SANITIZE_FFT=1 timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/san_mem.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/san_mem_dense.log 2>&1
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err
python bench.py --steps 100 --warmup 5 --N_r 20 --N_fm 128 --members-per-gpu 1024 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
python bench.py --steps 50 --warmup 5 --N_r 40 --N_fm 512 --members-per-gpu 512 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_under_ncu.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/prof_fft python tools/profile_step.py > gpurun_out/prof_fft.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/prof_jvp python tools/profile_step.py jvp > gpurun_out/prof_jvp.log 2>&1
tail -n 1 gpurun_out/smoke.log; tail -n 1 gpurun_out/pytest_gpu.log; tail -n 1 gpurun_out/san_mem.log; tail -n 1 gpurun_out/san_mem_dense.log; cat gpurun_out/bench_reference_arm.json | head -c 600; echo
python - <<'PY'
import json
for f in ("bench_1gpu","bench_cfg2","bench_cfg5"):
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "jvp", d["jvp"]["value"], d["jvp"].get("uncached"), "diag", d["with_diagnostics"]["value"], d["clocks"], d["gpu_launches"])
PY

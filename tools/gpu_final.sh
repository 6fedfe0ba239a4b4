# Round-end evidence on one GPU (gpurun -- 'bash tools/gpu_final.sh'): smoke, the whole GPU test suite, stage times, the
# ncu launch list of the bench command and three --set full captures (step, cached JVP, time loop with diagnostics),
# summarised on the box (tools/summarize_ncu.py); only the step's report itself travels back (64 MiB limit).
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 300 python tools/stage_times.py > gpurun_out/stages.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --parity-steps 0 --strong-members 0 > gpurun_out/bench_under_ncu.log 2>&1
for mode in step jvp diag; do
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o /tmp/prof_r02_$mode python tools/profile_step.py $mode > gpurun_out/prof_r02_$mode.log 2>&1
  python tools/summarize_ncu.py /tmp/prof_r02_$mode.ncu-rep gpurun_out/ncu_r02_$mode.txt > /dev/null 2>&1
done
cp /tmp/prof_r02_step.ncu-rep gpurun_out/
tail -n 1 gpurun_out/smoke.log; tail -n 3 gpurun_out/pytest_gpu.log; tail -n 7 gpurun_out/stages.log; ls -la gpurun_out/*.ncu-rep gpurun_out/ncu_r02_*

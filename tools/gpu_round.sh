# one GPU-box round trip: full bench at the headline shape, the config-2 / config-5 shapes, launch list
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
python bench.py --steps 100 --warmup 5 --N_r 20 --N_fm 128 --members-per-gpu 1024 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
python bench.py --steps 50 --warmup 5 --N_r 40 --N_fm 512 --members-per-gpu 512 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
SDDC_FFT=0 python bench.py --steps 50 --warmup 5 --N_r 40 --N_fm 512 --members-per-gpu 512 --no-cpu-baseline > gpurun_out/bench_cfg5_dense.json 2> gpurun_out/bench_cfg5_dense.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_under_ncu.log 2>&1
for f in bench_1gpu bench_cfg2 bench_cfg5 bench_cfg5_dense; do echo "== $f"; head -c 600 gpurun_out/$f.json; echo; tail -n 2 gpurun_out/$f.err; done

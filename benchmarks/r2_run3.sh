set -x
python benchmarks/timeline.py 8 6 8 > gpurun_out/r2c_timeline_s8_g8.json 2> gpurun_out/r2c_timeline.err; tail -c 300 gpurun_out/r2c_timeline.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 1 --streams 1 --group 8 --no-cpu-baseline > gpurun_out/r2c_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_rp_|k_msm_" -s 34 -c 17 -o gpurun_out/r2c_full python bench.py --steps 1 --warmup 0 --streams 1 --group 8 --no-cpu-baseline > gpurun_out/r2c_ncu_full.log 2>&1
ls -la gpurun_out | head -30

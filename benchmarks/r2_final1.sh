set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_reference.json 2> gpurun_out/r2_final_reference.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -c 300 gpurun_out/r2_final_bench.err
python -c "import json; d=json.loads(open('gpurun_out/r2_final_bench.json').read()); print('FINAL value', round(d['value']), 'e2e', round(d['e2e']['value']), d['clocks'], d['cpu_baseline'], d['roofline']['frac'], d['roofline']['int_pipe']['frac'])"
python bench.py --gpus 1 --steps 20 --warmup 5 --m 16 --batch 256 --no-cpu-baseline > gpurun_out/r2_final_bench_config3.json 2> gpurun_out/r2_final_bench_config3.err; python -c "import json; d=json.loads(open('gpurun_out/r2_final_bench_config3.json').read()); print('CFG3 value', round(d['value']), 'e2e', round(d['e2e']['value']))"
python benchmarks/timeline.py 8 6 8 > gpurun_out/r2_final_timeline.json 2> gpurun_out/r2_final_timeline.err
python benchmarks/reject_path.py > gpurun_out/r2_final_reject.json 2>/dev/null; cat gpurun_out/r2_final_reject.json
python benchmarks/secondary.py 12 > gpurun_out/r2_final_secondary.json 2> gpurun_out/r2_final_secondary.err; tail -3 gpurun_out/r2_final_secondary.err; head -c 1500 gpurun_out/r2_final_secondary.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 1 --streams 1 --group 8 --no-cpu-baseline > gpurun_out/r2_final_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_rp_|k_msm_" -s 28 -c 14 -o gpurun_out/r2_final_full python bench.py --steps 1 --warmup 0 --streams 1 --group 8 --no-cpu-baseline > gpurun_out/r2_final_ncu_full.log 2>&1
ls -la gpurun_out/r2_final*

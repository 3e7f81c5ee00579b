set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python benchmarks/reject_path.py > gpurun_out/r2e_reject.json 2> gpurun_out/r2e_reject.err; cat gpurun_out/r2e_reject.json; tail -3 gpurun_out/r2e_reject.err
for cfg in "8 8" "6 8"; do set -- $cfg; python bench.py --steps 20 --warmup 5 --streams $1 --group $2 --no-cpu-baseline > gpurun_out/r2e_s$1_g$2.json 2> gpurun_out/r2e_s$1_g$2.err; python -c "import sys,json; d=json.loads(open('gpurun_out/r2e_s$1_g$2.json').read()); print('S G', $1, $2, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['roofline']['per_kernel_ms_per_group'])"; done

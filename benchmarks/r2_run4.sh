set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for cfg in "8 8" "6 8" "12 4" "16 4"; do set -- $cfg; python bench.py --steps 20 --warmup 5 --streams $1 --group $2 --no-cpu-baseline > gpurun_out/r2d_s$1_g$2.json 2> gpurun_out/r2d_s$1_g$2.err; python -c "import sys,json; d=json.loads(open('gpurun_out/r2d_s$1_g$2.json').read()); print('S G', $1, $2, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['roofline']['per_kernel_ms_per_group'])"; done
python benchmarks/timeline.py 8 6 8 > gpurun_out/r2d_timeline_s8_g8.json 2> gpurun_out/r2d_timeline.err

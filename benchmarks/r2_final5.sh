set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_final5_bench.json 2> gpurun_out/r2_final5_bench.err; tail -c 300 gpurun_out/r2_final5_bench.err
python -c "import json; d=json.loads(open('gpurun_out/r2_final5_bench.json').read().strip().splitlines()[-1]); print('FINAL5 value', round(d['value']), 'e2e', round(d['e2e']['value']), d['clocks'])"

set -x
for rep in 1 2; do
unset BP_EXP_LEAN; python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2j_base_$rep.json 2> gpurun_out/r2j_base.err; python -c "import json; d=json.loads(open('gpurun_out/r2j_base_$rep.json').read()); print('VAR base value', round(d['value']), 'e2e', round(d['e2e']['value']))"
BP_EXP_LEAN=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2j_lean_$rep.json 2> gpurun_out/r2j_lean.err; python -c "import json; d=json.loads(open('gpurun_out/r2j_lean_$rep.json').read()); print('VAR lean value', round(d['value']), 'e2e', round(d['e2e']['value']), d['roofline']['per_kernel_ms_per_group'])"
done
BP_EXP_LEAN=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2

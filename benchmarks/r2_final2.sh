set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final2_bench.json 2> gpurun_out/r2_final2_bench.err; tail -c 300 gpurun_out/r2_final2_bench.err
python -c "import json; d=json.loads(open('gpurun_out/r2_final2_bench.json').read()); print('FINAL2 value', round(d['value']), 'e2e', round(d['e2e']['value']), d['diag'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['int_pipe'].get('ncu_fmaheavy'))"
python benchmarks/launch_cost.py > gpurun_out/r2_launch_cost.json 2>/dev/null; cat gpurun_out/r2_launch_cost.json
python benchmarks/prover_batch.py 1 64 1024 > gpurun_out/r2_final2_prover.json 2>/dev/null; cat gpurun_out/r2_final2_prover.json
python benchmarks/r1cs_shuffle.py 32769 > gpurun_out/r2_final2_r1cs.json 2>/dev/null; cat gpurun_out/r2_final2_r1cs.json

set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for rep in 1 2; do
python bench.py --steps 20 --warmup 5 --streams 8 --group 8 --no-cpu-baseline > gpurun_out/r2g_base_$rep.json 2> gpurun_out/r2g_base.err; python -c "import json; d=json.loads(open('gpurun_out/r2g_base_$rep.json').read()); print('BASE value', round(d['value']), 'e2e', round(d['e2e']['value']))"
BPMSM_LIB_EXPERIMENT=$PWD/bulletproofs_b200/libbpmsm_alu.so python bench.py --steps 20 --warmup 5 --streams 8 --group 8 --no-cpu-baseline > gpurun_out/r2g_alu_$rep.json 2> gpurun_out/r2g_alu.err; python -c "import json; d=json.loads(open('gpurun_out/r2g_alu_$rep.json').read()); print('ALU value', round(d['value']), 'e2e', round(d['e2e']['value']), d['roofline']['per_kernel_ms_per_group'])"
done
BPMSM_LIB_EXPERIMENT=$PWD/bulletproofs_b200/libbpmsm_alu.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "field or msm_matches or golden or launch_groups" 2>&1 | tail -3
for lg in 10 16 20; do python bench.py --workload msm --lg $lg --steps 10 --warmup 3 > gpurun_out/r2g_msm_$lg.json 2> gpurun_out/r2g_msm_$lg.err; tail -2 gpurun_out/r2g_msm_$lg.err; python -c "import json; d=json.loads(open('gpurun_out/r2g_msm_$lg.json').read()); print('MSM lg', $lg, 'value', round(d['value']), 'with_decompress', round(d['with_decompress']['value']), 'e2e', round(d['e2e']['value']), 'e2e_comp', round(d['e2e_compressed']['value']), d['parity'], d['roofline']['int_pipe']['frac'], d['roofline']['per_kernel_ms_per_call'])"; done
python benchmarks/prover_batch.py 1 64 1024 > gpurun_out/r2g_prover.json 2> gpurun_out/r2g_prover.err; cat gpurun_out/r2g_prover.json
python benchmarks/r1cs_shuffle.py 32769 > gpurun_out/r2g_r1cs.json 2> gpurun_out/r2g_r1cs.err; cat gpurun_out/r2g_r1cs.json

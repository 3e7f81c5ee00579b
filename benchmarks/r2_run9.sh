set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for rep in 1 2; do for v in "" noTR noDEC noBOTH; do
  if [ -z "$v" ]; then unset BPMSM_LIB_EXPERIMENT; tag=stage; else export BPMSM_LIB_EXPERIMENT=$PWD/bulletproofs_b200/libbpmsm_$v.so; tag=$v; fi
  python bench.py --steps 20 --warmup 5 --streams 8 --group 8 --no-cpu-baseline > gpurun_out/r2i_${tag}_$rep.json 2> gpurun_out/r2i_$tag.err; python -c "import json; d=json.loads(open('gpurun_out/r2i_${tag}_$rep.json').read()); print('VARIANT $tag value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done; done
unset BPMSM_LIB_EXPERIMENT
python benchmarks/reject_path.py > gpurun_out/r2i_reject.json 2> gpurun_out/r2i_reject.err; cat gpurun_out/r2i_reject.json
for lg in 10 14 16 18 20; do python bench.py --workload msm --lg $lg --steps 10 --warmup 3 > gpurun_out/r2i_msm_$lg.json 2> gpurun_out/r2i_msm_$lg.err; tail -2 gpurun_out/r2i_msm_$lg.err; python -c "import json; d=json.loads(open('gpurun_out/r2i_msm_$lg.json').read()); print('MSM lg', $lg, 'value', round(d['value']), 'with_decompress', round(d['with_decompress']['value']), 'e2e', round(d['e2e']['value']), 'e2e_comp', round(d['e2e_compressed']['value']), d['parity'], round(d['roofline']['int_pipe']['frac'],3), d['roofline']['per_kernel_ms_per_call'])"; done
for w in 13 16; do python bench.py --workload msm --lg 20 --window $w --steps 10 --warmup 3 > gpurun_out/r2i_msm_20_w$w.json 2> gpurun_out/r2i_msm_20_w$w.err; python -c "import json; d=json.loads(open('gpurun_out/r2i_msm_20_w$w.json').read()); print('MSM lg 20 window', $w, 'value', round(d['value']), d['roofline']['per_kernel_ms_per_call'])"; done
python benchmarks/prover_batch.py 1 64 1024 > gpurun_out/r2i_prover.json 2> gpurun_out/r2i_prover.err; cat gpurun_out/r2i_prover.json
python benchmarks/r1cs_shuffle.py 32769 > gpurun_out/r2i_r1cs.json 2> gpurun_out/r2i_r1cs.err; cat gpurun_out/r2i_r1cs.json

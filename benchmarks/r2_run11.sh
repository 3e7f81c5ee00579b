set -x
for rep in 1 2; do
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2k_base_$rep.json 2>/dev/null; python -c "import json; d=json.loads(open('gpurun_out/r2k_base_$rep.json').read()); print('OCC base value', round(d['value']), 'e2e', round(d['e2e']['value']))"
for v in d6a5 d5a6 d6a6; do python benchmarks/ab_lib.py benchmarks/ab/$v.so --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2k_${v}_$rep.json 2>/dev/null; python -c "import json; d=json.loads(open('gpurun_out/r2k_${v}_$rep.json').read()); print('OCC $v value', round(d['value']), 'e2e', round(d['e2e']['value']), d['roofline']['per_kernel_ms_per_group']['k_rp_decompress'], d['roofline']['per_kernel_ms_per_group']['k_msm_accumulate'])"; done
done
bash benchmarks/r2_scale.sh 1 sweep

set -x
timeout 300 python -m pytest tests/test_gpu_fp64.py tests/test_gpu_parity.py -m gpu -x -q -k "fp64 or field or decompress or launch_groups" 2>&1 | tail -4
timeout 120 python benchmarks/fp64_ladder.py 20 > gpurun_out/r2_fp64_ladder.json 2> gpurun_out/r2_fp64_ladder.err; cat gpurun_out/r2_fp64_ladder.json; tail -3 gpurun_out/r2_fp64_ladder.err
timeout 120 python benchmarks/fp64_ladder.py 20 bulletproofs_b200/libbpmsm_fdint.so > gpurun_out/r2_fp64_ladder_int.json 2>> gpurun_out/r2_fp64_ladder.err; cat gpurun_out/r2_fp64_ladder_int.json
for sh in 0 8 4 6; do
  timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --fp64-share $sh > gpurun_out/r2_fp64_s$sh.json 2> gpurun_out/r2_fp64_s$sh.err
  python -c "import json; d=json.loads(open('gpurun_out/r2_fp64_s$sh.json').read().strip().splitlines()[-1]); print('FP64 share', $sh, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['clocks'])"
done
for sh in 8 6; do
  timeout 200 python benchmarks/ab_lib.py bulletproofs_b200/libbpmsm_fdint.so --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --fp64-share $sh > gpurun_out/r2_fp64_int_s$sh.json 2> gpurun_out/r2_fp64_int_s$sh.err
  python -c "import json; d=json.loads(open('gpurun_out/r2_fp64_int_s$sh.json').read().strip().splitlines()[-1]); print('FP64 intfinish share', $sh, 'value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done
for sh in 0 8; do
  timeout 200 python bench.py --workload msm --lg 20 --steps 5 --warmup 3 --fp64-share $sh > gpurun_out/r2_fp64_msm_s$sh.json 2> gpurun_out/r2_fp64_msm_s$sh.err
  python -c "import json; d=json.loads(open('gpurun_out/r2_fp64_msm_s$sh.json').read().strip().splitlines()[-1]); print('FP64 msm share', $sh, 'value', round(d['value']), 'with_decompress', d.get('with_decompress'), 'e2e', d['e2e'])" 2>&1 | cut -c1-600
done

set -x
for rep in 1 2; do
python bench.py --steps 20 --warmup 5 --streams 8 --group 8 --no-cpu-baseline > gpurun_out/r2h_stage_$rep.json 2> gpurun_out/r2h_stage.err; python -c "import json; d=json.loads(open('gpurun_out/r2h_stage_$rep.json').read()); print('STAGE value', round(d['value']), 'e2e', round(d['e2e']['value']))"
BPMSM_LIB_EXPERIMENT=$PWD/bulletproofs_b200/libbpmsm_nostage.so python bench.py --steps 20 --warmup 5 --streams 8 --group 8 --no-cpu-baseline > gpurun_out/r2h_nostage_$rep.json 2> gpurun_out/r2h_nostage.err; python -c "import json; d=json.loads(open('gpurun_out/r2h_nostage_$rep.json').read()); print('NOSTAGE value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done
ncu --set full --clock-control none --import-source on -k regex:"k_msm_accumulate|k_msm_scatter|k_msm_count" -s 6 -c 3 -o gpurun_out/r2h_msm20 python bench.py --workload msm --lg 20 --steps 1 --warmup 0 > gpurun_out/r2h_ncu_msm20.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_msm_accumulate" -s 2 -c 1 -o gpurun_out/r2h_msm16 python bench.py --workload msm --lg 16 --steps 1 --warmup 0 > gpurun_out/r2h_ncu_msm16.log 2>&1
ls -la gpurun_out/*.ncu-rep

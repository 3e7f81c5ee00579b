# usage: bash benchmarks/r2_scale.sh N [sweep]   (inside gpurun --gpus N)
N=$1
mkdir -p gpurun_out
run() { if [ "$N" = "1" ]; then python bench.py --gpus 1 "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; fi; }
run --impl reference --steps 3 --warmup 1 > gpurun_out/r2_scale_ref_n$N.json 2> gpurun_out/r2_scale_ref_n$N.err
run --steps 20 --warmup 5 > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err
tail -c 600 gpurun_out/r2_scale_n$N.err
python -c "import json; d=json.loads(open('gpurun_out/r2_scale_n$N.json').read().strip().splitlines()[-1]); print('SCALE N', $N, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['clocks'], d['diag'])"
SIZES=""
if [ "$2" = "sweep" ]; then SIZES="10 11 12 13 14 15 16 17 18 19 20"; fi
if [ "$2" = "sweep3" ]; then SIZES="10 16 20"; fi
if [ -n "$SIZES" ]; then
  for lg in $SIZES; do
    run --workload msm --lg $lg --steps 10 --warmup 3 > gpurun_out/r2_msm_n${N}_lg$lg.json 2> gpurun_out/r2_msm_n${N}_lg$lg.err
    python -c "import json; d=json.loads(open('gpurun_out/r2_msm_n${N}_lg$lg.json').read().strip().splitlines()[-1]); print('MSM N', $N, 'lg', $lg, 'value', round(d['value']), 'with_decompress', round(d['with_decompress']['value']), 'e2e', round(d['e2e']['value']), 'e2e_comp', round(d['e2e_compressed']['value']), d['parity'], round(d['roofline']['int_pipe']['frac'],3))" || tail -3 gpurun_out/r2_msm_n${N}_lg$lg.err
  done
fi

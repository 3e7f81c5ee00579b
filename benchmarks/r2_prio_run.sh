set -x
for pol in 0 1 2 1 0; do
  BP_GRAPH_PRIORITY=$pol timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_prio_p$pol.json 2> gpurun_out/r2_prio_p$pol.err
  python -c "import json; d=json.loads(open('gpurun_out/r2_prio_p$pol.json').read().strip().splitlines()[-1]); print('PRIO policy', $pol, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  cp gpurun_out/r2_prio_p$pol.json gpurun_out/r2_prio_p${pol}_last.json
done
BP_GRAPH_PRIORITY=1 timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
BP_GRAPH_PRIORITY=1 python -c "import __graft_entry__ as g; g.smoke()"

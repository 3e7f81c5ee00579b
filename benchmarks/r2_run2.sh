set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for cfg in "4 8" "6 8" "8 8" "8 4" "12 4" "16 4" "12 2" "16 2" "10 8" "6 16"; do set -- $cfg; python bench.py --steps 20 --warmup 5 --streams $1 --group $2 --no-cpu-baseline > gpurun_out/r2b_s$1_g$2.json 2> gpurun_out/r2b_s$1_g$2.err; python -c "import sys,json; d=json.loads(open('gpurun_out/r2b_s$1_g$2.json').read()); print('S G', $1, $2, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['diag']['host_issue_ms_per_step'], d['diag']['host_issue_ms_per_step_e2e'])"; done

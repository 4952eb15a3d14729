set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_final_tests.log
tail -4 gpurun_out/r2_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_final_smoke.log 2>&1; tail -2 gpurun_out/r2_final_smoke.log
timeout 600 python bench.py > gpurun_out/r2_final_c1.json 2> gpurun_out/r2_final_c1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err
timeout 300 python bench.py --config 4 --steps 10 --warmup 3 > gpurun_out/r2_final_c4.json 2> gpurun_out/r2_final_c4.err
python - <<'PY'
import json
for f in ['c1','ref','c4']:
    try:
        d=json.loads(open(f'gpurun_out/r2_final_{f}.json').read().strip().splitlines()[-1])
        print(f, d.get('value'), d.get('ms_per_step'), (d.get('roofline') or {}).get('frac'), d.get('e2e',{}).get('value'), d.get('clocks'), (d.get('cpu_baseline') or {}).get('value'), (d.get('gpu_eager_baseline') or {}))
    except Exception as e:
        print(f, 'ERR', e)
PY

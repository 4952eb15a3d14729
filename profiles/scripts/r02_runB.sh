set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_entrypoints_gpu.py tests/test_optim_gpu.py -q -m gpu 2>&1 | tail -40 > gpurun_out/r2_runB_tests.log
tail -5 gpurun_out/r2_runB_tests.log
timeout 300 python bench.py --config 4 --steps 10 --warmup 3 > gpurun_out/r2_runB_c4.json 2> gpurun_out/r2_runB_c4.err
timeout 300 python bench.py --config 4 --steps 10 --warmup 3 --no-graph > gpurun_out/r2_runB_c4_eager.json 2> gpurun_out/r2_runB_c4_eager.err
timeout 300 python bench.py --config 1 --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_runB_c1.json 2> gpurun_out/r2_runB_c1.err
timeout 200 python bench.py --config 1 --batch 1 --steps 200 --warmup 20 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_runB_b1.json 2> gpurun_out/r2_runB_b1.err
timeout 200 python bench.py --config 1 --batch 8 --steps 100 --warmup 10 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_runB_b8.json 2> gpurun_out/r2_runB_b8.err
python - <<'PY'
import json
for f in ['c4','c4_eager','c1','b1','b8']:
    try:
        d=json.loads(open(f'gpurun_out/r2_runB_{f}.json').read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d.get('host_issue_ms_per_step'), (d.get('roofline') or {}).get('frac'), d.get('e2e',{}).get('value'))
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -3 gpurun_out/r2_runB_c4.err

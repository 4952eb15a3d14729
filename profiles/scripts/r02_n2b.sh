set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29521 bench.py --gpus 2 --config 4 --steps 10 --warmup 3 > gpurun_out/r2c_n2_c4.json 2> gpurun_out/r2c_n2_c4.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2c_n2_c4.json').read().strip().splitlines()[-1])
    print(d.get('value'), d.get('ms_per_step'), d.get('allreduce'), d['config'].get('cuda_graph'), d['config'].get('cuda_graph_error'), d.get('last_step'))
except Exception as e:
    print('ERR', e)
PY
tail -5 gpurun_out/r2c_n2_c4.err

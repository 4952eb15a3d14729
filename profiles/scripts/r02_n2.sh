set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 bench.py --gpus 2 --config 4 --steps 10 --warmup 3 > gpurun_out/r2b_n2_c4.json 2> gpurun_out/r2b_n2_c4.err
TORCH_NCCL_ASYNC_ERROR_HANDLING=0 ODB_TRAIN_GRAPH_COLLECTIVES=1 timeout 300 $TR --master-port 29512 bench.py --gpus 2 --config 4 --steps 10 --warmup 3 > gpurun_out/r2b_n2_c4_graph.json 2> gpurun_out/r2b_n2_c4_graph.err
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --config 1 --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2b_n2_c1.json 2> gpurun_out/r2b_n2_c1.err
python - <<'PY'
import json
for f in ['c4','c4_graph','c1']:
    try:
        d=json.loads(open(f'gpurun_out/r2b_n2_{f}.json').read().strip().splitlines()[-1])
        print(f, d.get('value'), d.get('ms_per_step'), d.get('allreduce'), d['config'].get('cuda_graph'), d.get('first_step'), d.get('last_step'))
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -5 gpurun_out/r2b_n2_c4_graph.err

nvidia-smi -L | head -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 300 --warmup 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=2', d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'], d['clocks'])"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_device" 2>&1 | tail -2
timeout 300 python tools/engine_bench.py --impl ours --seconds 3 --moves 2 --threads 16 --gpus 2 2>&1 | cut -c1-400

run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 5 --warmup 3 --no-side --shard ladder 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$2', '%.4e e2e %.4e' % (d['value'], d['e2e']['value']))"; }
run 29551 default
NCCL_MAX_NCHANNELS=4 run 29552 maxch4
NCCL_MAX_NCHANNELS=8 run 29553 maxch8
NCCL_MAX_NCHANNELS=2 run 29554 maxch2

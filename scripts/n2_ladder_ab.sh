run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 5 --warmup 3 --no-side $2 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$3', '%.4e e2e %.4e' % (d['value'], d['e2e']['value']))"; }
run 29521 "" walkers
run 29522 "--shard ladder" ladder_bulk
PTMCMC_NO_BULK_GROUP=1 run 29523 "--shard ladder" ladder_nobulk
run 29524 "--shard ladder" ladder_bulk
PTMCMC_NO_BULK_GROUP=1 run 29525 "--shard ladder" ladder_nobulk
run 29526 "" walkers

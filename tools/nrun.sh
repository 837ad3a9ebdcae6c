# usage: bash tools/nrun.sh N [extra bench args]   -> one bench line summary at N GPUs
N=${1:-2}; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 300 --warmup 5 "$@" 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], 'value', d['value'], 'ms/step', d['ms_per_step'], 'wall', d['wall_ms_per_step'], 'e2e', d['e2e']['value'])
"

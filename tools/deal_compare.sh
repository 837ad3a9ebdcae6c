#!/usr/bin/env bash
# N GPUs: the bench with the cost-aware deal of the views and with other flags (default: the round-robin deal), back to back on one box;
# per-rank times in the line.   tools/deal_compare.sh <N> "<extra flags of run 2>"
N=${1:-8}
i=0
for extra in "" "${2:---round-robin}"; do
  i=$((i+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 --no-configs --no-cpu-baseline $extra > gpurun_out/deal_n${N}_run$i.json 2> gpurun_out/deal_n${N}_run$i.err
  python - "gpurun_out/deal_n${N}_run$i.json" "$extra" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        j = json.loads(l)
        print(repr(sys.argv[2]), "value", j["value"], "ms/step", j["ms_per_step"], "parity", j["parity"]["visbuffer_exact"], "gather", j.get("gather_check", {}).get("matching"), "deal", j.get("deal"), "per_rank", j.get("per_rank"))
PY
done

#!/usr/bin/env bash
# N = 8: the bench with the cost-aware deal of the views and with the round-robin deal (v mod N), back to back on one box.
for pol in "" "--round-robin"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3 --no-configs --no-cpu-baseline $pol > gpurun_out/deal_n8_${pol:-cost}.json 2> gpurun_out/deal_n8_${pol:-cost}.err
  python - "gpurun_out/deal_n8_${pol:-cost}.json" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        j = json.loads(l)
        print(sys.argv[1], "value", j["value"], "ms/step", j["ms_per_step"], "e2e", j["e2e"]["value"], "parity", j["parity"]["visbuffer_exact"], j["parity"]["views_checked"], "gather", j.get("gather_check", {}).get("matching"), "deal", j.get("deal"))
PY
done

#!/usr/bin/env bash
# What limits the strong-scaled batch at N GPUs: the same bench with the composite exchange on (p2p), off (none) and with
# fewer / more views in flight; the line's `nvlink` object carries rank 0's NVLink byte counters over the timed region.
#   tools/n8_limiter.sh <N> <out-dir>
N=${1:-8}; out=${2:-gpurun_out}
run() { tag=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus "$N" --steps 50 --warmup 3 --no-configs --no-cpu-baseline "$@" > "$out/lim_n${N}_$tag.json" 2> "$out/lim_n${N}_$tag.err"
  python - "$out/lim_n${N}_$tag.json" "$tag" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        j = json.loads(l)
        print(sys.argv[2], "value", j["value"], "ms/step", j["ms_per_step"], "e2e", j["e2e"]["value"], "submit_us", j["host_submit_us_per_view"], "nvlink", j.get("nvlink"))
PY
}
nvidia-smi nvlink -gt d -i 0 | head -8
run p2p --gather p2p
run none --gather none
run nccl --gather nccl
run p2p_f2 --gather p2p --in-flight 2
run p2p_e2ediag --gather p2p --e2e-diag
grep -a "e2e-diag" "$out/lim_n${N}_p2p_e2ediag.err"

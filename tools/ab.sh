show() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[2], 'value', d['value'], 'ms', d['ms_per_step'], 'lat', d['latency_ms_per_frame'], {k:v['us'] for k,v in d['stages'].items()})" $1 "$2"; }
python -m pytest tests/test_resolve_gpu.py tests/test_debug_programs.py -m gpu -q 2>&1 | tail -1
for m in 1 2; do for f in 5 6 8; do
python bench.py --no-cpu-baseline --mesh-blocks $m --in-flight $f > gpurun_out/ab_tmp.json 2>/dev/null; show gpurun_out/ab_tmp.json "mesh=$m F=$f"
done; done

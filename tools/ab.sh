show() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[2], 'value', d['value'], 'ms', d['ms_per_step'], 'lat', d['latency_ms_per_frame'], {k:v['us'] for k,v in d['stages'].items()})" $1 "$2"; }
for m in 1 2; do for f in 4 5 6; do
SWRB_MESH_BLOCKS_PER_SM=$m python bench.py --no-cpu-baseline --in-flight $f > gpurun_out/ab_tmp.json 2>/dev/null; show gpurun_out/ab_tmp.json "mesh_blocks=$m F=$f"
done; done

show() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[2], 'value', d['value'], 'ms', d['ms_per_step'], 'lat', d['latency_ms_per_frame'], {k:v['us'] for k,v in d['stages'].items()})" $1 $2; }
python -m pytest tests/test_resolve_gpu.py tests/test_light_markers.py -m gpu -x -q 2>&1 | tail -3
for v in "" _mb5 _mb6; do
SWRB_LIB=$PWD/glimpsw_b200/libswrb$v.so python bench.py --no-cpu-baseline > gpurun_out/ab$v.json 2>/dev/null; show gpurun_out/ab$v.json "lib$v"
done
SWRB_LIB=$PWD/glimpsw_b200/libswrb_mb5.so python bench.py --no-cpu-baseline > gpurun_out/ab_r2.json 2>/dev/null; show gpurun_out/ab_r2.json "lib_mb5_again"

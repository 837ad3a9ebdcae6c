for a in 64 128 256; do for f in "3 2" "4 2" "3 4" "4 1"; do set -- $f; echo "== inline $a in-flight $1 mesh-blocks $2"; SWRB_INLINE_AREA=$a python bench.py --no-cpu-baseline --no-configs --steps 10 --in-flight $1 --mesh-blocks $2 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', d['value'], 'ms/view', d['ms_per_view'], 'latency', d['latency_ms_per_view'], 'e2e', d['e2e']['value'], 'parity', d['parity']['visbuffer_exact'])
"; done; done

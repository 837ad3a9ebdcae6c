# developer sweep: contexts in flight x mesh-kernel blocks per SM on the contract bench (no CPU baseline, no config block)
for f in "4 2" "4 3" "4 4" "5 2" "6 2" "6 3" "8 2"; do set -- $f; echo "== in-flight $1 mesh-blocks $2"; python bench.py --no-cpu-baseline --no-configs --steps 10 --in-flight $1 --mesh-blocks $2 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', d['value'], 'ms/view', d['ms_per_view'], 'latency', d['latency_ms_per_view'], 'e2e', d['e2e']['value'], 'host us/view', d['host_submit_us_per_view'], 'parity', d['parity']['visbuffer_exact'])
"; done

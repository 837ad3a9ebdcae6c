# one multi-GPU box: C5 views at N = 1, 2, 4, 8 and the contract bench at N = 4, 8
for n in 1 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n tools/c5_views.py 2>/dev/null | grep '^{' | tee gpurun_out/c5_n$n.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C5', d['n_gpus'], d['views_per_s'], 'views/s', d['ms_per_view_per_gpu'], 'ms/view/gpu', d['Gtri_per_s_submitted'], 'Gtri/s')"
done
for n in 4 8; do bash tools/nrun.sh $n --no-cpu-baseline; done

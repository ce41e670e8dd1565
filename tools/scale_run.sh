#!/bin/bash
# Runs bench.py at N = 1, 2, 4, 8 on one box (whatever GPUs are visible) and prints the key numbers.
# WORKLOAD=random1m bash tools/scale_run.sh 1 2 4 8   for the 1M-path 16384x16384 configuration.
W=${WORKLOAD:-headline}
for N in "$@"; do
  if [ "$N" = 1 ]; then CMD="python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --workload $W";
  else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --workload $W"; fi
  $CMD 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('N=%d value=%.3f Gseg/s ms/step=%.3f frames=%s e2e=%.3f (%.2f ms)' % (d['n_gpus'], d['value'], d['ms_per_step'], {k: round(v,3) for k,v in d['config']['ms_per_frame'].items()}, d['e2e']['value'], d['e2e']['ms_per_step']))"
done

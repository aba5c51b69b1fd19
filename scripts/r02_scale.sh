#!/bin/bash
# e2e / device scaling on one 8-GPU box: N = 8, 4, 2, 1 back to back (driver-style launch)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_8gpu.txt 2>&1
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 10 --warmup 3 --no-cpu --no-small --stream-steps 0 > gpurun_out/r02_scale_n$n.json 2> gpurun_out/r02_scale_n$n.err
done
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --no-small --stream-steps 0 > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err
python - <<PY
import json
for n in (1,2,4,8):
    try:
        d=json.loads(open(f'gpurun_out/r02_scale_n{n}.json').read().strip().splitlines()[-1])
        print(n, 'value %.0f'%d['value'], 'ms %.1f'%d['ms_per_step'], 'e2e %.0f (%.1f ms)'%(d['e2e']['value'], d['e2e']['ms_per_step']), 'dense %.0f'%d['e2e_dense_format']['value'], 'h2d %.1f GB/s'%d['h2d_only']['gb_per_s_per_gpu'])
    except Exception as e:
        print(n, 'failed', e)
PY

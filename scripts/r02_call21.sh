#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_runners.py -m gpu -q --timeout 300 2>&1 | tail -3
python - <<'PY'
import torch, time, numpy as np
from spgnn_b200 import runner, synth_device, ops
from spgnn_b200._lib import lib, ptr, stream
g = synth_device.make_batch(first_tree=0, count=4096, seed=1234, ragged=False).graph
hb = runner.host_batch_from_graph(g, packed=True)
bufs = runner._upload(hb, torch.device('cuda',0))
n_nodes, e_off, e_src, e_dst, f_row_off, f_mask, f_vals, fvs_out, labels = bufs
N = int(f_mask.shape[0])
fvs = ops.empty_padded(N, hb.fv_dim, torch.device('cuda',0))
L = lib()
def call(): L.unpack_rows(ptr(f_mask), ptr(f_vals), ptr(f_row_off), N, hb.fv_dim, ptr(fvs), fvs.stride(0), stream())
call(); torch.cuda.synchronize()
a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): call()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)/10
byt = 4.0*hb.f_nnz + 4.0*N*hb.fv_dim + N*(hb.fv_dim//8+8)
print('unpack_rows %.3f ms  %.0f GB/s' % (ms, byt/ms/1e6), 'equal', bool(torch.equal(fvs[:, :hb.fv_dim], g.ndata['fvs'])))
PY
timeout -k 5 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-small --stream-steps 0 --e2e-steps 12 > gpurun_out/e2e.json 2>gpurun_out/e2e.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/e2e.json').read().strip().splitlines()[-1])
print('step %.2f'%d['ms_per_step'], 'e2e %.2f ms %.0f g/s'%(d['e2e']['ms_per_step'], d['e2e']['value']))
PY

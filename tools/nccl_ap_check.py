#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/nccl_ap_check.py -- every rank accumulates a DIFFERENT number of TP/FP records on its
device; the device tail (tfe.gather_tp_fp_records over NCCL + tfe.average_precision_records) must equal the host path
(tfe.gather_tp_fp + precision_recall + average_precision_voc07 / _voc12) on every rank: AP07 bit for bit, AP12 to 1e-12."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import ron_tensorflow_b200.tf_extended as tfe  # noqa: E402

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
dist.init_process_group('nccl')
C = 21
rng = np.random.Generator(np.random.PCG64(100 + rank))
state = tfe.TpFpDeviceState(C, capacity=1 << 18)
for u in range(1 + 2 * rank):                                  # rank r: 1 + 2 r updates of r-dependent size
    B, M = 2 + 3 * rank + u, 200
    sc = np.round(rng.uniform(0, 1, size=(B, C - 1, M)), 2).astype(np.float32)
    sc[rng.uniform(size=sc.shape) < 0.4] = 0.
    tp = rng.uniform(size=sc.shape) < 0.2
    fp = (~tp) & (rng.uniform(size=sc.shape) < 0.7)
    ng = rng.integers(0, 9, size=(B, C - 1)).astype(np.int64)
    state.update(*[torch.from_numpy(x).cuda() for x in (ng, tp, fp, sc)])
ap07, ap12 = state.average_precision()
merged = tfe.gather_tp_fp(state, C)
n = 0
for c in range(1, C):
    p_, r_ = tfe.precision_recall(*merged[c].value())
    n += merged[c].scores.shape[0]
    assert ap07[c] == tfe.average_precision_voc07(p_, r_), (rank, c, ap07[c])
    assert abs(ap12[c] - tfe.average_precision_voc12(p_, r_)) <= 1e-12, (rank, c)
rec, _ = tfe.gather_tp_fp_records(state)
assert int((rec != tfe.PAD_RECORD).sum()) == n
if world > 1:
    assert rec.numel() > n, 'ranks hold different record counts: the gathered rows must carry padding'
dist.barrier()
if rank == 0:
    print('nccl_ap_check OK: %d ranks, %d records, mAP07 %.6f' % (world, n, float(np.mean(list(ap07.values())))))
dist.destroy_process_group()

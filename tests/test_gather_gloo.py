"""The path's only collective (tfe.gather_tp_fp: all_gather of per-class TP/FP records +
all_reduce of GT counts) on CPU with the gloo backend, world size 2: the merged state must be
what a single process would have accumulated for the rank-major concatenation of the shards."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NUM_CLASSES = 5


def _records(rank):
    rng = np.random.Generator(np.random.PCG64(1234 + rank))
    out = {}
    for c in range(1, NUM_CLASSES):
        n = int(rng.integers(0, 40)) if not (rank == 1 and c == 2) else 0       # one empty shard
        scores = rng.uniform(0., 1., n).astype(np.float32)
        scores[rng.uniform(size=n) < 0.1] = 0.                                 # zero scores are dropped (:169-171)
        tp = rng.uniform(size=n) < 0.4
        fp = (~tp) & (rng.uniform(size=n) < 0.7)
        out[c] = (int(rng.integers(0, 9)), tp, fp, scores)
    return out


def _accumulate(tfe, shards):
    state = None
    for rec in shards:
        _, state = tfe.streaming_tp_fp_arrays({c: np.array([rec[c][0]]) for c in rec}, {c: rec[c][1] for c in rec},
                                              {c: rec[c][2] for c in rec}, {c: rec[c][3] for c in rec}, state=state)
    return state


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import ron_tensorflow_b200.tf_extended as tfe
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    state = _accumulate(tfe, [_records(rank)])
    merged = tfe.gather_tp_fp(state, NUM_CLASSES)
    res = {c: (merged[c].n_gt, merged[c].scores.copy(), merged[c].tp.copy(), merged[c].fp.copy()) for c in merged}
    aps = {}
    for c in merged:
        p, r = tfe.precision_recall(*merged[c].value())
        aps[c] = (tfe.average_precision_voc07(p, r), tfe.average_precision_voc12(p, r))
    # the other collective of the path: the detections themselves (rank-major = a contiguous image split)
    rng = np.random.Generator(np.random.PCG64(77 + rank))
    ds = torch.from_numpy(rng.uniform(size=(3, NUM_CLASSES - 1, 6)).astype(np.float32))
    db = torch.from_numpy(rng.uniform(size=(3, NUM_CLASSES - 1, 6, 4)).astype(np.float32))
    gs, gb = tfe.gather_detections(ds, db)
    dd_s, dd_b = tfe.gather_detections({c: ds[:, c - 1] for c in range(1, NUM_CLASSES)}, {c: db[:, c - 1] for c in range(1, NUM_CLASSES)})
    q.put((rank, res, aps, gs.numpy(), gb.numpy(), {c: v.numpy() for c, v in dd_s.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_tp_fp_world2_gloo():
    import torch.multiprocessing as mp
    import ron_tensorflow_b200.tf_extended as tfe
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = _accumulate(tfe, [_records(0), _records(1)])
    want_s, want_b = [], []
    for r in range(2):
        rng = np.random.Generator(np.random.PCG64(77 + r))
        want_s.append(rng.uniform(size=(3, NUM_CLASSES - 1, 6)).astype(np.float32))
        want_b.append(rng.uniform(size=(3, NUM_CLASSES - 1, 6, 4)).astype(np.float32))
    want_s, want_b = np.concatenate(want_s), np.concatenate(want_b)
    for rank, res, aps, gs, gb, dd_s in got:
        assert np.array_equal(gs, want_s) and np.array_equal(gb, want_b)
        for c in range(1, NUM_CLASSES):
            assert np.array_equal(dd_s[c], want_s[:, c - 1])
        for c in range(1, NUM_CLASSES):
            n_gt, scores, tp, fp = res[c]
            assert n_gt == single[c].n_gt
            assert np.array_equal(scores, single[c].scores)
            assert np.array_equal(tp, single[c].tp) and np.array_equal(fp, single[c].fp)
            p, r = tfe.precision_recall(*single[c].value())
            assert aps[c] == (tfe.average_precision_voc07(p, r), tfe.average_precision_voc12(p, r))


def test_gather_is_identity_without_process_group():
    import ron_tensorflow_b200.tf_extended as tfe
    state = _accumulate(tfe, [_records(0)])
    assert tfe.gather_tp_fp(state, NUM_CLASSES) is state
    x, y = torch.zeros(2, 3, 4), torch.zeros(2, 3, 4, 4)
    gx, gy = tfe.gather_detections(x, y)
    assert gx is x and gy is y

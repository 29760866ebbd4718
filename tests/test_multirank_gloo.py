"""CPU, world_size 2 over gloo: the host-side logic of the N>1 path -- shard assignment and the
one exchange step of the evaluation (all-gather of both sets, row blocks of the CD matrices,
all-gather of the row blocks).  The CD matrix itself is injected from the oracle here (the CUDA
kernel cannot run on CPU); on the GPU the same code path runs with NCCL and the CUDA kernel."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pcd_b200
from oracle import pointdiff_oracle as O
from conftest import same_set_metrics


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 8, 8192, 8191):
        for world in (1, 2, 4, 8):
            spans = [pcd_b200.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        pcd_b200.shard_range(8, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, G, R, out, tile=512):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gs, gc = pcd_b200.shard_range(G.shape[0], rank, world)
        rs, rc = pcd_b200.shard_range(R.shape[0], rank, world)
        res = pcd_b200.evaluate_sets(G[gs:gs + gc], R[rs:rs + rc], matrix_fn=O.chamfer_matrix, tile=tile)
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_evaluate_sets_world2_equals_single_process():
    g = torch.Generator().manual_seed(3)
    G = torch.randn(6, 64, 3, generator=g) * torch.rand(6, 1, 3, generator=g)
    R = torch.randn(6, 64, 3, generator=g) * torch.rand(6, 1, 3, generator=g)
    want = O.set_metrics_from_matrices(O.chamfer_matrix(G, R), O.chamfer_matrix(G, G), O.chamfer_matrix(R, R))
    single = pcd_b200.evaluate_sets(G, R, matrix_fn=O.chamfer_matrix)
    assert same_set_metrics(single, want)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), G, R, out), nprocs=2, join=True)
    assert dict(out[0]) == dict(out[1]) and same_set_metrics(dict(out[0]), want)


def test_evaluate_sets_uneven_shards_and_tiles():
    """7 generated / 5 reference clouds on 2 ranks (shard_range leaves 4 + 3 and 3 + 2): the all-gathers must cope with
    unequal blocks, and a tile smaller than the sets exercises the round-robin triangle schedule and the MIN all-reduces."""
    g = torch.Generator().manual_seed(4)
    G = torch.randn(7, 48, 3, generator=g) * torch.rand(7, 1, 3, generator=g)
    R = torch.randn(5, 48, 3, generator=g) * torch.rand(5, 1, 3, generator=g)
    want = O.set_metrics_from_matrices(O.chamfer_matrix(G, R), O.chamfer_matrix(G, G), O.chamfer_matrix(R, R))
    assert same_set_metrics(pcd_b200.evaluate_sets(G, R, matrix_fn=O.chamfer_matrix, tile=2), want)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), G, R, out, 2), nprocs=2, join=True)
    assert dict(out[0]) == dict(out[1]) and same_set_metrics(dict(out[0]), want)

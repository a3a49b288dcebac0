"""Sequence sharding + logits all-gather, world_size 2 on the gloo backend (CPU)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from esme import parallel as P
from oracle import esm_oracle as O


def test_partition_is_a_balanced_cover():
    lens = O.synthetic_lengths(50000, seed=2)
    for world in (1, 2, 4, 8):
        owned = P.partition_sequences(lens, world)
        flat = sorted(i for o in owned for i in o)
        assert flat == list(range(len(lens)))
        assert P.imbalance(lens, owned) < 1.12
    assert P.partition_sequences([5], 4) == [[0], [], [], []]


def test_shard_plan_row_maps_agree():
    """The two row maps of a ShardPlan describe the same permutation: perm (packed row -> row of the rank-major padded
    all-gather buffer, used by esmk_allgather_logits) and dest_rows (this rank's row -> packed row, used by the peer
    path's direct stores); every packed row is owned by exactly one rank."""
    lens = [40, 7, 300, 2, 128, 129, 64, 5]
    tokens, cu, _ = O.synthetic_batch(lens, seed=9)
    for world in (1, 2, 3, 8, 16):                    # 16 ranks: some own nothing
        plans = [P.ShardPlan(tokens, cu, world, r, 64, 'cpu') for r in range(world)]
        T = int(cu[-1])
        seen = torch.zeros(T, dtype=torch.int64)
        for r, pl in enumerate(plans):
            assert pl.dest_rows.dtype == torch.int32 and pl.dest_rows.numel() == pl.tokens.numel()
            seen[pl.dest_rows.long()] += 1
            assert torch.equal(tokens[pl.dest_rows.long()], pl.tokens)
            # row i of rank r sits at r * t_max + i in the gathered buffer
            assert torch.equal(pl.perm[pl.dest_rows.long()], torch.arange(pl.tokens.numel()) + r * pl.t_max)
            assert torch.equal(pl.perm, plans[0].perm) and pl.t_max == plans[0].t_max
        assert bool((seen == 1).all())
    P.close_comms()                                   # nothing cached on the CPU path: a no-op


def _fake_forward(tokens, cu_lens, max_len):
    """Deterministic stand-in for the per-rank GPU forward: depends on the token, its position
    inside its own sequence and that sequence's length (so wrong sharding / ordering shows)."""
    pos = O.positions_from_cu_lens(cu_lens.cpu())
    lens = (cu_lens[1:] - cu_lens[:-1]).long()
    seq_len = torch.repeat_interleave(lens, lens)
    base = tokens.float() * 3 + pos.float() * 0.5 + seq_len.float() * 0.25
    return (base[:, None] + torch.arange(5)[None, :]).to(torch.bfloat16)


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lens = [40, 7, 300, 2, 128, 129, 64]
        tokens, cu, _ = O.synthetic_batch(lens, seed=9)
        out = P.sharded_forward(_fake_forward, tokens, cu, width=5, embed_dim=64, device='cpu')
        want = _fake_forward(tokens, cu, max(lens))
        ret[rank] = bool(torch.equal(out, want))
    finally:
        dist.destroy_process_group()


def test_sharded_forward_world2_gloo():
    world = 2
    port = 29500 + os.getpid() % 2000
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}

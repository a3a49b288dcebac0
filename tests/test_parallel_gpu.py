"""Real multi-GPU check (needs >= 2 GPUs, one process per GPU over NCCL): the sequence-sharded forward with the
logits all-gather behind the C ABI -- direct NVLink peer stores (esmk_peer_allgather_logits), ncclAllGather + row
gather (esmk_allgather_logits), or torch.distributed -- must equal the single-GPU forward bit for bit
(batch-composition invariance makes the per-rank forwards identical to the un-sharded one)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret, collective):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), ESMK_COLLECTIVE=collective)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        import esme
        from esme import parallel, synthetic
        model = esme.ESM.from_pretrained(f'{GOLDEN}/esm2_tiny.safetensors', device=dev)
        lens = [300, 131, 66, 2, 129, 257, 40]
        tokens, cu, max_len = synthetic.synthetic_batch(lens, seed=21)
        got = parallel.model_sharded_forward(model, tokens, cu, kind='logits')
        logp = parallel.model_sharded_forward(model, tokens, cu, kind='log_prob')
        want = model(tokens.to(dev), (cu.to(dev), max_len))
        ok = bool(torch.equal(got, want)) and bool(torch.equal(
            logp, model.predict_log_prob(tokens.to(dev), (cu.to(dev), max_len))))
        # repeated calls alternate the two window buffers; a batch whose sequences all land on one rank leaves the other
        # rank with zero rows
        for seed, ls in ((22, [50, 60, 70, 80]), (23, [200]), (24, lens)):
            t2, c2, m2 = synthetic.synthetic_batch(ls, seed=seed)
            plan = parallel.ShardPlan(t2, c2, world, rank, model.embed_dim, dev)
            out = model(plan.tokens.to(dev), (plan.cu_lens.to(dev), plan.max_len)) if plan.tokens.numel() else \
                torch.zeros(0, 33, dtype=torch.bfloat16, device=dev)
            ok = ok and bool(torch.equal(plan.gather(out), model(t2.to(dev), (c2.to(dev), m2))))
            used = plan.collective
        want_impl = {'peer': 'esmk_peer_allgather_logits', 'nccl': 'esmk_allgather_logits', 'torch': 'torch.distributed'}
        ret[rank] = ok and used.startswith(want_impl[collective])
        parallel.close_comms()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
@pytest.mark.parametrize('collective', ['peer', 'nccl', 'torch'])
def test_sharded_forward_equals_single_gpu(collective):
    world = 2
    port = 29700 + os.getpid() % 2000 + ['peer', 'nccl', 'torch'].index(collective)
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, ret, collective), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}

"""Sequence-sharded multi-GPU forward (one process per GPU, torch.distributed).

The reference has no inference-time multi-GPU path; sequences are
independent (attention never crosses cu_lens boundaries), so a packed batch is
split by whole sequences, every rank runs the single-GPU forward on its share
with replicated weights, and ONE collective -- an all-gather of the per-rank
logits -- restores the original token order on every rank.  No collective is
needed inside the forward.
"""
import ctypes as C
import os
import sys
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


class EsmkComm:
    """NCCL communicator owned by libesmk (esmk_comm_t): the collective of the path runs behind the C ABI
    (esmk_allgather_logits = ncclAllGather + packed-order row gather).  torch.distributed is only used to hand the
    128-byte NCCL unique id from rank 0 to the other ranks."""

    def __init__(self, group=None, device=None):
        from . import _lib as L
        self.L = L
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device(device)
        buf = (C.c_char * 128)()
        if self.rank == 0:
            L.check(L.lib.esmk_comm_unique_id(buf), 'esmk_comm_unique_id')
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        idbuf = C.create_string_buffer(box[0], 128)
        self.handle = L.c_void_p()
        # NCCL prints its version banner on stdout when a communicator is created outside torch (NCCL_DEBUG=WARN /
        # VERSION environments): keep stdout clean for callers that parse it (bench.py prints one JSON line)
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            with torch.cuda.device(self.device):
                L.check(L.lib.esmk_comm_create(C.byref(self.handle), self.world, self.rank, idbuf), 'esmk_comm_create')
        finally:
            os.dup2(saved, 1)
            os.close(saved)

    def allgather_rows(self, local: torch.Tensor, perm: torch.Tensor, gathered: torch.Tensor, out: torch.Tensor):
        L = self.L
        t_max, V = local.shape
        with torch.cuda.device(self.device):
            L.check(L.lib.esmk_allgather_logits(self.handle, local.data_ptr(), t_max, V, perm.data_ptr(), out.shape[0],
                                                gathered.data_ptr(), out.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream), 'esmk_allgather_logits')
        return out

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                self.L.lib.esmk_comm_destroy(self.handle)
        except Exception:
            pass


def sequence_cost(length: int, embed_dim: int, ffn_factor: float = 24.0) -> float:
    """FLOP model per sequence: L * (8 + ffn_factor) * D^2 linear work + 4 * D * L^2 attention
    (per layer; the layer count is a common factor).  ffn_factor = 16 (ESM2) .. 6F/D (ESMC)."""
    return length * ffn_factor * embed_dim * embed_dim + 4.0 * embed_dim * length * length


def partition_sequences(lengths: Sequence[int], world_size: int, embed_dim: int = 1280) -> List[List[int]]:
    """Longest-processing-time greedy: sequences sorted by cost, each given to the least
    loaded rank.  Returns, per rank, the sorted list of sequence indices it owns."""
    order = sorted(range(len(lengths)), key=lambda i: -sequence_cost(lengths[i], embed_dim))
    load = [0.0] * world_size
    owned: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        owned[r].append(i)
        load[r] += sequence_cost(lengths[i], embed_dim)
    return [sorted(o) for o in owned]


def imbalance(lengths: Sequence[int], owned: List[List[int]], embed_dim: int = 1280) -> float:
    """max / mean of the per-rank modelled cost."""
    loads = [sum(sequence_cost(lengths[i], embed_dim) for i in o) for o in owned]
    mean = sum(loads) / len(loads)
    return max(loads) / mean if mean > 0 else 1.0


def take_sequences(tokens: torch.Tensor, cu_lens: torch.Tensor, idx: Sequence[int]):
    """Packed sub-batch made of sequences `idx`: (tokens, cu_lens int32, max_len, token_index)."""
    cu = cu_lens.tolist()
    if len(idx) == 0:
        e = torch.empty(0, dtype=torch.int64)
        return tokens[:0], torch.zeros(1, dtype=torch.int32), 0, e
    spans = [torch.arange(cu[i], cu[i + 1], dtype=torch.int64) for i in idx]
    token_index = torch.cat(spans)
    lens = torch.tensor([cu[i + 1] - cu[i] for i in idx], dtype=torch.int32)
    sub_cu = torch.zeros(len(idx) + 1, dtype=torch.int32)
    sub_cu[1:] = torch.cumsum(lens, 0)
    return tokens[token_index], sub_cu, int(lens.max()), token_index


class ShardPlan:
    """Partition of one packed batch over the ranks of `group`, built once per batch on the host: which sequences
    each rank owns, this rank's packed sub-batch, and the row permutation that turns the all-gathered, per-rank
    padded logits back into the original packed order.  `gather()` is the single collective of the path."""

    def __init__(self, tokens: torch.Tensor, cu_lens: torch.Tensor, world: int, rank: int, embed_dim: int = 1280,
                 device=None):
        self.world, self.rank, self.device = world, rank, device
        self.lengths = (cu_lens[1:] - cu_lens[:-1]).tolist()
        self.owned = partition_sequences(self.lengths, world, embed_dim)
        self.shares = [take_sequences(tokens, cu_lens, o) for o in self.owned]
        self.t_max = max(int(s[3].numel()) for s in self.shares)
        self.T = int(cu_lens[-1])
        self.tokens, self.cu_lens, self.max_len, self.token_index = self.shares[rank]
        self.imbalance = imbalance(self.lengths, self.owned, embed_dim)
        # perm[t] = row of the gathered [world * t_max, width] buffer holding packed token t
        perm = torch.empty(self.T, dtype=torch.int64)
        for r, s in enumerate(self.shares):
            perm[s[3]] = torch.arange(s[3].numel(), dtype=torch.int64) + r * self.t_max
        self.perm = perm.to(device)
        self._local = self._gathered = None
        self.comm: Optional[EsmkComm] = None      # created on the first CUDA gather (collective over the group)
        self.collective = None                    # which implementation ran the collective (reported by bench.py)

    def gather(self, out: torch.Tensor, group=None) -> torch.Tensor:
        """out: this rank's [T_r, width] result -> [T, width] in the original packed order, on every rank."""
        width = out.shape[1]
        if self._local is None or self._local.shape[1] != width or self._local.dtype != out.dtype:
            self._local = torch.zeros(self.t_max, width, dtype=out.dtype, device=self.device)
            self._gathered = torch.empty(self.world * self.t_max, width, dtype=out.dtype, device=self.device)
        self._local[:out.shape[0]] = out
        if out.is_cuda and out.dtype == torch.bfloat16 and os.environ.get('ESMK_COLLECTIVE', 'esmk') != 'torch':
            # the only collective of the path, behind the C ABI: ncclAllGather + packed-order row gather in libesmk
            if self.comm is None and self.collective is None:
                try:
                    self.comm = EsmkComm(group, self.device)
                    self.collective = 'esmk_allgather_logits (ncclAllGather + row gather in libesmk, C ABI)'
                except Exception as e:      # e.g. no loadable libnccl.so.2: every rank fails alike (same image)
                    print(f'esme.parallel: libesmk communicator unavailable ({e}); using torch.distributed NCCL', file=sys.stderr)
                    self.collective = 'torch.distributed.all_gather_into_tensor (NCCL) + index_select'
            if self.comm is not None:
                result = torch.empty(self.T, width, dtype=out.dtype, device=self.device)
                return self.comm.allgather_rows(self._local, self.perm, self._gathered, result)
        if self.collective is None:
            self.collective = 'torch.distributed.all_gather_into_tensor + index_select'
        dist.all_gather_into_tensor(self._gathered, self._local, group=group)      # (CPU / gloo tests, ESMK_COLLECTIVE=torch)
        return self._gathered.index_select(0, self.perm)


def sharded_forward(fn: Callable[[torch.Tensor, torch.Tensor, int], torch.Tensor], tokens: torch.Tensor,
                    cu_lens: torch.Tensor, width: int, embed_dim: int = 1280, group=None,
                    device=None, dtype=torch.bfloat16) -> torch.Tensor:
    """Run `fn(tokens_r, cu_lens_r, max_len_r) -> [T_r, width]` on this rank's share of
    the batch (host tensors in, device tensors to fn) and all-gather the results into
    the original packed order: returns [T, width] on every rank.

    tokens / cu_lens are the FULL batch as host tensors, identical on every rank."""
    plan = ShardPlan(tokens, cu_lens, dist.get_world_size(group), dist.get_rank(group), embed_dim, device)
    if plan.tokens.numel() > 0:
        out = fn(plan.tokens.to(device), plan.cu_lens.to(device), plan.max_len)
    else:
        out = torch.zeros(0, width, dtype=dtype, device=device)
    return plan.gather(out, group)


def model_sharded_forward(model, tokens, cu_lens, kind: str = 'logits', group=None):
    """Convenience wrapper for an esme model living on this rank's GPU."""
    dev = next(model.parameters()).device
    fns = {'logits': lambda t, c, m: model(t, (c, m)),
           'log_prob': lambda t, c, m: model.predict_log_prob(t, (c, m))}
    return sharded_forward(fns[kind], tokens, cu_lens, model.lm_head.final.out_features,
                           model.embed_dim, group, dev)

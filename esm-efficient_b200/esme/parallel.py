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
    """Communicator owned by libesmk (esmk_comm_t): the collective of the path runs behind the C ABI, either as direct
    NVLink peer stores into every rank's window (esmk_peer_allgather_logits, after enable_peer) or as
    esmk_allgather_logits = ncclAllGather + packed-order row gather.  torch.distributed is only used to hand the
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

    peer_bytes = 0          # size of one buffer of the peer window (0: not enabled)

    def enable_peer(self, buffer_bytes: int) -> bool:
        """Collective: map every rank's window (CUDA IPC).  False where that is impossible -- on every rank alike."""
        L = self.L
        with torch.cuda.device(self.device):
            rc = L.lib.esmk_comm_enable_peer(self.handle, int(buffer_bytes))
        if rc != 0:
            self.peer_error = L.lib.esmk_last_error().decode()
            return False
        self.peer_bytes = int(buffer_bytes)
        return True

    def peer_allgather_rows(self, local: torch.Tensor, dest_rows: torch.Tensor, out: torch.Tensor):
        """local [T_r, V] bf16 (this rank's rows), dest_rows int32 [T_r] (their packed rows) -> out [T, V]."""
        L = self.L
        rows, V = local.shape
        with torch.cuda.device(self.device):
            L.check(L.lib.esmk_peer_allgather_logits(self.handle, local.data_ptr() if rows else None, rows, V,
                                                     dest_rows.data_ptr() if rows else None, out.shape[0], out.data_ptr(),
                                                     torch.cuda.current_stream().cuda_stream), 'esmk_peer_allgather_logits')
        return out

    def allgather_rows(self, local: torch.Tensor, perm: torch.Tensor, gathered: torch.Tensor, out: torch.Tensor):
        L = self.L
        t_max, V = local.shape
        with torch.cuda.device(self.device):
            L.check(L.lib.esmk_allgather_logits(self.handle, local.data_ptr(), t_max, V, perm.data_ptr(), out.shape[0],
                                                gathered.data_ptr(), out.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream), 'esmk_allgather_logits')
        return out

    def close(self, collective: bool = False):
        """Release the communicator (NCCL comm, peer windows and mappings).  Call it on every rank while all ranks are
        still alive -- close_comms() does, behind a barrier -- rather than leaving it to interpreter shutdown."""
        handle, self.handle = getattr(self, 'handle', None), None
        if handle:
            with torch.cuda.device(self.device):
                torch.cuda.synchronize()
                if collective and self.peer_bytes:
                    self.L.lib.esmk_comm_disable_peer(handle)      # unmap peers, meet, free the window (in that order)
                self.L.lib.esmk_comm_destroy(handle)
        self.peer_bytes = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_COMMS = {}


def close_comms(group=None):
    """Collective: destroy the cached libesmk communicators in step on all ranks (device idle, nobody mid-collective)."""
    if not _COMMS:
        return
    if dist.is_initialized():
        dist.barrier(group)
    for comm in list(_COMMS.values()):
        comm.close(collective=True)
    _COMMS.clear()
    if dist.is_initialized():
        dist.barrier(group)


def get_comm(group, device) -> EsmkComm:
    """One libesmk communicator per (process group, device), created on first use (collective over the group)."""
    key = (id(group) if group is not None else 0, torch.device(device).index)
    if key not in _COMMS:
        _COMMS[key] = EsmkComm(group, device)
    return _COMMS[key]


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
        self.dest_rows = self.token_index.to(torch.int32).to(device)      # packed row of each of this rank's rows
        self._local = self._gathered = None
        self.comm: Optional[EsmkComm] = None      # created on the first CUDA gather (collective over the group)
        self.collective = None                    # which implementation ran the collective (reported by bench.py)

    PEER = 'esmk_peer_allgather_logits (direct NVLink peer stores into every rank\'s window + flag handshake in libesmk, C ABI; no NCCL on the data path)'
    NCCL = 'esmk_allgather_logits (ncclAllGather + row gather in libesmk, C ABI)'

    def gather(self, out: torch.Tensor, group=None) -> torch.Tensor:
        """out: this rank's [T_r, width] result -> [T, width] in the original packed order, on every rank.
        ESMK_COLLECTIVE = nccl (default) | peer (NVLink peer stores, opt-in) | torch selects the implementation."""
        width = out.shape[1]
        want = os.environ.get('ESMK_COLLECTIVE', 'nccl')
        if out.is_cuda and out.dtype == torch.bfloat16 and want != 'torch':
            if self.comm is None and self.collective is None:
                try:
                    self.comm = get_comm(group, self.device)
                except Exception as e:      # e.g. no loadable libnccl.so.2: every rank fails alike (same image)
                    print(f'esme.parallel: libesmk communicator unavailable ({e}); using torch.distributed NCCL', file=sys.stderr)
                    self.collective = 'torch.distributed.all_gather_into_tensor (NCCL) + index_select'
            if self.comm is not None:
                result = torch.empty(self.T, width, dtype=out.dtype, device=self.device)
                need = self.T * width * out.element_size()
                if want == 'peer' and self.comm.peer_bytes == 0 and not getattr(self.comm, 'peer_failed', False):
                    # window sized for batches up to 4x this one (every rank sees the same T: same decision everywhere)
                    if not self.comm.enable_peer(max(4 * need, 64 << 20)):
                        self.comm.peer_failed = True
                        print(f'esme.parallel: NVLink peer window unavailable ({self.comm.peer_error}); using ncclAllGather',
                              file=sys.stderr)
                if want == 'peer' and self.comm.peer_bytes >= need:
                    self.collective = self.PEER
                    return self.comm.peer_allgather_rows(out.contiguous(), self.dest_rows, result)
                self.collective = self.NCCL
                self._stage(out, width)
                return self.comm.allgather_rows(self._local, self.perm, self._gathered, result)
        if self.collective is None:
            self.collective = 'torch.distributed.all_gather_into_tensor + index_select'
        self._stage(out, width)
        dist.all_gather_into_tensor(self._gathered, self._local, group=group)      # (CPU / gloo tests, ESMK_COLLECTIVE=torch)
        return self._gathered.index_select(0, self.perm)

    def _stage(self, out: torch.Tensor, width: int):
        if self._local is None or self._local.shape[1] != width or self._local.dtype != out.dtype:
            self._local = torch.zeros(self.t_max, width, dtype=out.dtype, device=self.device)
            self._gathered = torch.empty(self.world * self.t_max, width, dtype=out.dtype, device=self.device)
        self._local[:out.shape[0]] = out


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

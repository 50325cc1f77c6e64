"""Sharded flat index: one process per GPU (torch.distributed), rows partitioned by contiguous label
ranges, queries replicated, one all-gather of the per-shard top-k (score, label) lists and a k-way
merge on every rank. The reference has no counterpart — its flat index is a single
in-process scan (algorithms/brute_force/brute_force.h:242-291); the merge order is the reference's
reply order, ascending (score, label).

torch is plumbing here: device buffers for the collective and NCCL itself. The scan and the merge are
libvsgpu kernels reached through the C-ABI (vsgpu_topk_device / vsgpu_merge_topk_device).
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from . import capi

_gpu_lib = None


def _vsgpu():
    global _gpu_lib
    if _gpu_lib is None:
        capi.lib()
        G = C.CDLL(os.path.join(capi.HERE, "libvsgpu.so"))
        vp, sz = C.c_void_p, C.c_size_t
        G.vsgpu_topk_device.argtypes = [vp, vp, sz, sz, sz, C.c_uint, vp, vp, vp]
        G.vsgpu_store_sync.argtypes = [vp]
        G.vsgpu_topk_device_begin.argtypes = [vp, vp, sz, sz, sz, C.c_uint, vp, vp, vp, C.c_uint, C.c_uint, vp]
        G.vsgpu_topk_device_next.argtypes = [vp, vp]
        G.vsgpu_topk_device_finish.argtypes = [vp, vp]
        G.vsgpu_topk_rounds.restype = sz
        G.vsgpu_topk_rounds.argtypes = [sz, sz, C.c_uint]
        G.vsgpu_store_stream.restype = vp
        G.vsgpu_store_stream.argtypes = [vp]
        G.vsgpu_store_set_stream.argtypes = [vp, vp]
        G.vsgpu_merge_topk_device.argtypes = [C.c_int, vp, C.c_int, sz, sz, sz, vp, vp, vp, vp]
        G.vsgpu_pack_topk_device.argtypes = [vp, sz, sz, vp, vp, vp]
        G.vsgpu_merge_packed_device.argtypes = [C.c_int, vp, sz, sz, sz, vp, vp, vp, vp, vp]
        G.vsgpu_packed_hit_bytes.restype = sz
        G.vsgpu_last_error.restype = C.c_char_p
        G.vsgpu_last_stats.argtypes = [vp, vp]
        _gpu_lib = G
    return _gpu_lib


class Stats(C.Structure):
    _fields_ = [("path", C.c_uint32), ("kernel_launches", C.c_uint32), ("candidates", C.c_uint64),
                ("fallback_queries", C.c_uint32), ("scan_ms", C.c_float), ("total_ms", C.c_float)]


def shard_bounds(n_total, world_size, rank):
    """Contiguous, balanced row range [lo, hi) of `rank`."""
    base, rem = divmod(n_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def merge_topk_host(scores, labels, k):
    """Reference merge semantics on the host (numpy): scores/labels [parts, nq, k] -> [nq, k] by
    ascending (score, label); padded entries carry label -1 / UINT64_MAX. Used by the CPU (gloo)
    tests of the collective plumbing and as the checker of the device merge kernel."""
    parts, nq, kk = scores.shape
    s = np.transpose(scores, (1, 0, 2)).reshape(nq, parts * kk)
    l = np.transpose(labels, (1, 0, 2)).reshape(nq, parts * kk).astype(np.uint64)
    out_s = np.full((nq, k), np.nan, dtype=scores.dtype)
    out_l = np.full((nq, k), np.iinfo(np.uint64).max, dtype=np.uint64)
    for q in range(nq):
        valid = l[q] != np.iinfo(np.uint64).max
        order = np.lexsort((l[q][valid], s[q][valid]))[:k]
        out_s[q, :len(order)] = s[q][valid][order]
        out_l[q, :len(order)] = l[q][valid][order]
    return out_s, out_l


class ShardedFlatIndex:
    """Rank-local shard + collective. Every rank calls every method (SPMD)."""

    def __init__(self, params, n_total=None, group=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = params
        self.f64 = params.type == capi.VecSimType_FLOAT64
        self.local = capi.BFIndex(params)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._store = None
        self._bufs = {}
        self._stream = None
        self._any_pin = None
        self._last = None
        self.redone_steps = 0
        self.share_bounds = os.environ.get("VSGPU_SHARE_BOUNDS", "1") != "0"   # A/B switch
        self.n_total = n_total          # rows over all shards; agreed on lazily after the rows changed (see _total)
        self._total_stale = n_total is None
        self.exchanges = 0

    def close(self):
        # torch's caching allocator ties a block to the stream it was allocated on: every buffer of this index is
        # allocated on the default stream (see _buf) and released before the store — and with it the stream — goes away
        if self._stream is not None:
            self._stream.synchronize()
        self._bufs.clear()
        self._any_pin = None
        self._stream = None
        torch.cuda.synchronize(self.device)
        self.local.close()

    # ---- ingest: the caller hands each rank its own rows (already sharded) ----
    def add_vectors(self, blobs, labels):
        self._total_stale = True
        return self.local.add_vectors(blobs, labels=labels)

    def add_device_rows(self, tensor, first_label):
        assert tensor.is_cuda and tensor.is_contiguous()
        torch.cuda.current_stream(self.device).synchronize()   # the rows were produced on the caller's stream
        self._total_stale = True
        return self.local.add_device_rows(tensor.data_ptr(), tensor.stride(0) * tensor.element_size(), tensor.shape[0],
                                          first_label)

    def _total(self):
        """Rows over all shards. Every rank calls every method (SPMD), so all ranks find the count stale at the same call
        and join the one all-reduce that refreshes it (a host wait, once after ingestion — not on the query path)."""
        if self._total_stale:
            if self.world > 1:
                t = torch.tensor([self.local.index_size()], dtype=torch.int64, device=self.device)
                dist.all_reduce(t, group=self.group)
                self.n_total = int(t.item())
            else:
                self.n_total = self.local.index_size()
            self._total_stale = False
        return self.n_total

    def store(self):
        if self._store is None:
            self._store = self.local.device_store()
        return self._store

    def _buf(self, name, shape, dtype):
        key = (name, tuple(shape), dtype)
        if key not in self._bufs:
            with torch.cuda.stream(torch.cuda.default_stream(self.device)):   # never on the store's (external) stream
                self._bufs[key] = torch.empty(shape, dtype=dtype, device=self.device)
        return self._bufs[key]

    def stream(self):
        """The store's CUDA stream as a torch stream. Every step runs with it as torch's current stream, so the H2D copy
        of the queries, the scan (enqueued by libvsgpu on that stream), NCCL and the merge are ordered by the stream
        itself: no host synchronisation anywhere on the path."""
        if self._stream is None:
            # a stream from torch's own pool, handed to the store: torch's allocators tie pinned and device blocks to the
            # streams that used them and touch those streams again when the blocks are freed — possibly after this index
            # is gone. Pool streams live as long as the process; a stream the store owned would not.
            self._stream = torch.cuda.Stream(device=self.device)
            rc = _vsgpu().vsgpu_store_set_stream(self.store(), C.c_void_p(self._stream.cuda_stream))
            if rc != 0:
                raise RuntimeError("vsgpu_store_set_stream: " + _vsgpu().vsgpu_last_error().decode())
        return self._stream

    def local_topk_device(self, q_dev, k, flags=0):
        """q_dev: [nq, blob] processed queries on this GPU -> (scores [nq,k], labels [nq,k]) device tensors.
        Enqueued on the store's stream (the caller orders its own stream against it, see topk_device)."""
        G = _vsgpu()
        nq = q_dev.shape[0]
        sdt = torch.float64 if self.f64 else torch.float32
        scores = self._buf("ls", (nq, k), sdt)
        labels = self._buf("ll", (nq, k), torch.int64)
        if self.world > 1 and not self.f64 and self.share_bounds and self._total() >= self.world:
            # phased: every shard runs the same number of coarse phases and after each ONE small all-reduce(max) makes the
            # shards' bounds global — the ceil(k / world)-th best of every shard bounds the k-th best overall — so the next
            # phase admits, and the re-rank scores, a shard's share of ONE candidate band instead of a band of its own (the
            # per-shard phases and the re-rank did not shrink with the shard: SURVEY §8e). The phase count
            # comes from n_total / world, which all ranks know, not from the local row count.
            rounds = int(G.vsgpu_topk_rounds((self.n_total + self.world - 1) // self.world, k, self.world))
            bounds = self._buf("bd", (2 * nq,), torch.float32)
            rc = G.vsgpu_topk_device_begin(self.store(), q_dev.data_ptr(), nq, q_dev.stride(0) * q_dev.element_size(), k, flags,
                                           labels.data_ptr(), scores.data_ptr(), None, self.world, rounds, bounds.data_ptr())
            if rc != 0:
                raise RuntimeError("vsgpu_topk_device_begin: " + G.vsgpu_last_error().decode())
            for _ in range(rounds - 1):
                dist.all_reduce(bounds, op=dist.ReduceOp.MAX, group=self.group)
                rc = G.vsgpu_topk_device_next(self.store(), bounds.data_ptr())
                if rc != 0:
                    raise RuntimeError("vsgpu_topk_device_next: " + G.vsgpu_last_error().decode())
            dist.all_reduce(bounds, op=dist.ReduceOp.MAX, group=self.group)
            rc = G.vsgpu_topk_device_finish(self.store(), bounds.data_ptr())
            if rc != 0:
                raise RuntimeError("vsgpu_topk_device_finish: " + G.vsgpu_last_error().decode())
            self.exchanges = rounds
            return scores, labels
        rc = G.vsgpu_topk_device(self.store(), q_dev.data_ptr(), nq, q_dev.stride(0) * q_dev.element_size(), k, flags,
                                 labels.data_ptr(), scores.data_ptr(), None)
        if rc != 0:
            raise RuntimeError("vsgpu_topk_device: " + G.vsgpu_last_error().decode())
        return scores, labels

    def topk_device(self, q_dev, k, flags=0):
        """Global top-k on every rank: local scan -> ONE all-gather of packed (label, score, flag) hits -> device merge.
        Everything is enqueued on the store's stream; results are final after finish()."""
        caller = torch.cuda.current_stream(self.device)
        st = self.stream()
        st.wait_stream(caller)                      # q_dev may have been produced on the caller's stream
        with torch.cuda.stream(st):
            out = self._topk_enqueue(q_dev, k, flags)
        caller.wait_stream(st)
        return out

    def _topk_enqueue(self, q_dev, k, flags):
        G = _vsgpu()
        scores, labels = self.local_topk_device(q_dev, k, flags)
        nq = q_dev.shape[0]
        self._last = (q_dev, k, nq)
        if self.world == 1:
            return scores, labels
        if self.f64:  # fp64 indexes never take the tensor path: no flags to carry, scores do not fit the packed hit
            all_s = self._buf("as", (self.world, nq, k), scores.dtype)
            all_l = self._buf("al", (self.world, nq, k), torch.int64)
            dist.all_gather_into_tensor(all_s, scores, group=self.group)
            dist.all_gather_into_tensor(all_l, labels, group=self.group)
            out_s = self._buf("os", (nq, k), scores.dtype)
            out_l = self._buf("ol", (nq, k), torch.int64)
            rc = G.vsgpu_merge_topk_device(self.device.index, C.c_void_p(self.stream().cuda_stream), 1, self.world, nq, k,
                                           all_s.data_ptr(), all_l.data_ptr(), out_s.data_ptr(), out_l.data_ptr())
            if rc != 0:
                raise RuntimeError("vsgpu_merge_topk_device: " + G.vsgpu_last_error().decode())
            return out_s, out_l
        return self._gather_merge(scores, labels, nq, k)

    def _gather_merge(self, scores, labels, nq, k):
        G = _vsgpu()
        hb = G.vsgpu_packed_hit_bytes()
        mine = self._buf("pk", (nq, k, hb), torch.uint8)
        rc = G.vsgpu_pack_topk_device(self.store(), nq, k, scores.data_ptr(), labels.data_ptr(), mine.data_ptr())
        if rc != 0:
            raise RuntimeError("vsgpu_pack_topk_device: " + G.vsgpu_last_error().decode())
        everyone = self._buf("pa", (self.world, nq, k, hb), torch.uint8)
        dist.all_gather_into_tensor(everyone, mine, group=self.group)
        out_s = self._buf("os", (nq, k), torch.float32)
        out_l = self._buf("ol", (nq, k), torch.int64)
        any_flag = self._buf("af", (1,), torch.int32)
        any_flag.zero_()
        rc = G.vsgpu_merge_packed_device(self.device.index, C.c_void_p(self.stream().cuda_stream), self.world, nq, k,
                                         everyone.data_ptr(), out_s.data_ptr(), out_l.data_ptr(), None, any_flag.data_ptr())
        if rc != 0:
            raise RuntimeError("vsgpu_merge_packed_device: " + G.vsgpu_last_error().decode())
        if self._any_pin is None:
            self._any_pin = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._any_pin.copy_(any_flag, non_blocking=True)
        return out_s, out_l

    def finish(self):
        """Wait for the last topk_device and make its result final: when some shard flagged a query (its tensor-path
        candidate buffer overflowed — adversarial ties), every rank sees the flag in the gathered hits; the shards redo
        their flagged queries on the exact path (vsgpu_store_sync) and the lists are gathered and merged once more.
        Returns the (scores, labels) device tensors of the step."""
        G = _vsgpu()
        st = self.stream()
        st.synchronize()
        redo = self.world > 1 and not self.f64 and self._any_pin is not None and int(self._any_pin[0]) != 0
        rc = G.vsgpu_store_sync(self.store())       # resolves this shard's own overflowed queries (no-op otherwise)
        if rc != 0:
            raise RuntimeError("vsgpu_store_sync: " + G.vsgpu_last_error().decode())
        if redo:
            _, k, nq = self._last
            with torch.cuda.stream(st):
                self._gather_merge(self._buf("ls", (nq, k), torch.float32), self._buf("ll", (nq, k), torch.int64), nq, k)
            st.synchronize()
            self.redone_steps += 1
        if self.world == 1:
            _, k, nq = self._last
            return self._buf("ls", (nq, k), torch.float64 if self.f64 else torch.float32), self._buf("ll", (nq, k), torch.int64)
        _, k, nq = self._last
        return self._buf("os", (nq, k), torch.float64 if self.f64 else torch.float32), self._buf("ol", (nq, k), torch.int64)

    def knn_batch(self, queries, k, flags=0, out_labels=None, out_scores=None):
        """Host queries (processed blobs, numpy [nq, blob] or a pinned uint8 tensor) -> host (labels int64 [nq,k], scores
        float64 [nq,k]). One H2D copy, the sharded search, one D2H copy per output into pinned buffers, one sync at the
        end; everything in between is ordered by the store's stream. out_labels (int64 / uint64 [nq,k]) and out_scores
        (float64 [nq,k]), when given, receive the reply in one pass each instead of fresh arrays."""
        if isinstance(queries, torch.Tensor):
            q = queries
        else:
            q = torch.from_numpy(np.ascontiguousarray(queries).view(np.uint8).reshape(queries.shape[0], -1))
        nq = q.shape[0]
        st = self.stream()
        st.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(st):
            q_dev = self._buf("qd", tuple(q.shape), torch.uint8)
            q_dev.copy_(q, non_blocking=True)
            self._topk_enqueue(q_dev, k, flags)
        out_s, out_l = self.finish()
        key = ("pin", nq, k, out_s.dtype)
        if key not in self._bufs:
            self._bufs[key] = (torch.empty((nq, k), dtype=out_s.dtype).pin_memory(), torch.empty((nq, k), dtype=torch.int64).pin_memory())
        pin_s, pin_l = self._bufs[key]
        with torch.cuda.stream(st):
            pin_s.copy_(out_s, non_blocking=True)
            pin_l.copy_(out_l, non_blocking=True)
        st.synchronize()
        if out_labels is not None and out_scores is not None:
            np.copyto(out_labels.view(np.int64), pin_l.numpy())
            np.copyto(out_scores, pin_s.numpy())                      # widens fp32 -> fp64 on the way
            return out_labels, out_scores
        return pin_l.numpy().copy(), pin_s.numpy().astype(np.float64)  # the pinned buffers are reused by the next call

    # ---- range query and batch iterator across shards: variable-length per-shard replies ----
    def range_query(self, query, radius, order=capi.BY_SCORE):
        """Raw query blob -> (labels int64 [n], scores float64 [n]) over all shards, ordered like the reference's
        rangeQuery + sort wrapper (vec_sim_index.h:246-252): by (score, label) or by label."""
        l, s = self.local.range_query(query, radius, order=capi.BY_SCORE)
        return merge_varlen(gather_varlen(s[0], l[0], self.group, self._coll_device()), order)

    def create_batch_iterator(self, query):
        return ShardedBatchIterator(self.local.create_batch_iterator(query), self.group, self._coll_device())

    def _coll_device(self):
        return self.device if dist.is_initialized() and dist.get_backend(self.group) == "nccl" else None

    def last_stats(self):
        st = Stats()
        _vsgpu().vsgpu_last_stats(self.store(), C.byref(st))
        return {f[0]: getattr(st, f[0]) for f in Stats._fields_}


def gather_merge_host(local_scores, local_labels, k, group=None):
    """The collective step with host tensors (gloo): all-gather per-shard lists, merge with the
    reference order. Mirrors ShardedFlatIndex.topk_device for CPU tests of the N>1 plumbing."""
    world = dist.get_world_size(group)
    s = torch.from_numpy(np.ascontiguousarray(local_scores))
    l = torch.from_numpy(np.ascontiguousarray(local_labels).view(np.int64))
    all_s = [torch.empty_like(s) for _ in range(world)]
    all_l = [torch.empty_like(l) for _ in range(world)]
    dist.all_gather(all_s, s, group=group)
    dist.all_gather(all_l, l, group=group)
    S = np.stack([t.numpy() for t in all_s])
    L = np.stack([t.numpy().view(np.uint64) for t in all_l])
    return merge_topk_host(S, L, k)


# ---- variable-length collectives (range replies, batch-iterator tails) -------------------------------------------------
def gather_varlen(scores, labels, group=None, device=None):
    """All-gather one (scores float64 [m_r], labels [m_r]) list per rank: counts first, then one padded all-gather.
    -> [(scores, labels uint64)] indexed by rank, identical on every rank. `device`: CUDA device for NCCL groups (the
    collective needs device tensors), None for host backends (gloo)."""
    scores = np.ascontiguousarray(scores, dtype=np.float64)
    labels = np.ascontiguousarray(labels).astype(np.uint64)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [(scores, labels)]
    world = dist.get_world_size(group)
    dev = device if device is not None else torch.device("cpu")
    cnt = torch.tensor([scores.size], dtype=torch.int64, device=dev)
    counts = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, cnt, group=group)
    counts = counts.cpu().numpy()
    width = int(counts.max())
    if width == 0:
        return [(np.zeros(0), np.zeros(0, dtype=np.uint64)) for _ in range(world)]
    pack = np.zeros((2, width), dtype=np.int64)  # scores travel as their bit patterns: one collective for both
    pack[0, :scores.size] = scores.view(np.int64)
    pack[1, :labels.size] = labels.view(np.int64)
    mine = torch.from_numpy(pack.reshape(-1)).to(dev)
    everything = torch.empty(world * 2 * width, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(everything, mine, group=group)
    everything = everything.cpu().numpy().reshape(world, 2, width)
    return [(everything[r, 0, :counts[r]].view(np.float64).copy(), everything[r, 1, :counts[r]].view(np.uint64).copy())
            for r in range(world)]


def merge_varlen(parts, order=capi.BY_SCORE):
    """Concatenate per-shard lists and order them as one reply: ascending (score, label), or ascending label (BY_ID).
    Shards hold disjoint labels, so there is nothing to de-duplicate."""
    s = np.concatenate([p[0] for p in parts]) if parts else np.zeros(0)
    l = np.concatenate([p[1] for p in parts]) if parts else np.zeros(0, dtype=np.uint64)
    idx = np.argsort(l, kind="stable") if order == capi.BY_ID else np.lexsort((l, s))
    return l[idx].astype(np.int64), s[idx]


class ShardedBatchIterator:
    """VecSimBatchIterator over all shards: every Next(n) returns the n globally best results not returned yet, in the
    order a single index over all rows would (bf_batch_iterator.h:59-214: ascending (score, label), n at a time).
    Each rank keeps the replicated list of what every shard has fetched but not yet returned; a call tops every shard's
    list up to n entries (one local Next + one variable-length all-gather), merges, and returns the first n."""

    def __init__(self, local_iterator, group=None, device=None):
        self.it, self.group, self.device = local_iterator, group, device
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._clear()

    def _clear(self):
        self.pending = [(np.zeros(0), np.zeros(0, dtype=np.uint64)) for _ in range(self.world)]
        self.depleted = [False] * self.world
        self.started = False

    def get_next_results(self, n, order=capi.BY_SCORE):
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        self.started = True
        want = max(n - self.pending[rank][0].size, 0)
        if want and self.it.has_next():
            l, s = self.it.get_next_results(want, order=capi.BY_SCORE)
            tail = (s[0], l[0].astype(np.uint64))
        else:
            tail = (np.zeros(0), np.zeros(0, dtype=np.uint64))
        # the depleted flag rides along as one extra entry (score NaN, label 1 = depleted)
        flag = (np.array([np.nan]), np.array([0 if self.it.has_next() else 1], dtype=np.uint64))
        got = gather_varlen(np.concatenate([tail[0], flag[0]]), np.concatenate([tail[1], flag[1]]), self.group, self.device)
        for r, (s, l) in enumerate(got):
            self.depleted[r] = bool(l[-1])
            self.pending[r] = (np.concatenate([self.pending[r][0], s[:-1]]), np.concatenate([self.pending[r][1], l[:-1]]))
        s = np.concatenate([p[0] for p in self.pending])
        l = np.concatenate([p[1] for p in self.pending])
        src = np.concatenate([np.full(p[0].size, r) for r, p in enumerate(self.pending)]).astype(np.int64)
        idx = np.lexsort((l, s))[:n]
        # a shard's entries leave in its own order (each list is ascending), so dropping counts from the front is exact
        for r in range(self.world):
            used = int((src[idx] == r).sum())
            self.pending[r] = (self.pending[r][0][used:], self.pending[r][1][used:])
        out_l, out_s = l[idx].astype(np.int64), s[idx]
        if order == capi.BY_ID:
            o = np.argsort(out_l, kind="stable")
            out_l, out_s = out_l[o], out_s[o]
        return out_l.reshape(1, -1), out_s.reshape(1, -1)

    def has_next(self):
        """False once every shard is depleted and nothing fetched is left (known after a Next, like the reference's
        tiered iterator: hnsw_tiered.h:1095-1106); before the first Next it is the local iterator's answer."""
        if not self.started:
            return self.it.has_next()
        return not (all(self.depleted) and all(p[0].size == 0 for p in self.pending))

    def reset(self):
        self.it.reset()
        self._clear()

    def close(self):
        self.it.close()

"""Multi-GPU counting: one process per GPU, key-range ownership, one exchange step.

The reference has no multi-device path (SURVEY.md 2.2 row C*); this is the one place
the path shards.  Each rank counts its own shard of reads into a
sorted key-unique run FIRST (local pre-aggregation: a key crosses NVLink once per
rank, with its count, instead of once per occurrence), cuts the run at the owners'
key-range splitters (contiguous slices, no partition kernel needed), exchanges the
slices with one all-to-all-v over NCCL, and merges the P slices it received with the
merge-path kernel.  Rank r then holds the final records of key range r; the global
artefact is the concatenation of the ranks' runs in rank order.

torch.distributed is plumbing only: it moves bytes between ranks.
"""
import numpy as np


# ------------------------------------------------------------------ ownership
def range_splitters(world, words):
    """P-1 ascending splitter keys cutting the key space into P equal-width ranges
    of the most significant word.  Shape [P-1, words] uint64."""
    sp = np.zeros((max(world - 1, 0), words), dtype=np.uint64)
    for i in range(1, world):
        sp[i - 1, 0] = np.uint64((i << 64) // world)
    return sp


def owner_of(keys_hi, world):
    """Owner rank of keys given their most significant word (host-side model, used by tests)."""
    sp = range_splitters(world, 1)[:, 0]
    return np.searchsorted(sp, np.asarray(keys_hi, dtype=np.uint64), side="right")


def slice_offsets_host(keys, splitters):
    """Host model of Run.split / kc_run_split: lower-bound offsets of the splitters in a
    sorted [n, W] uint64 key array -> P+1 offsets.  (CPU tests; the product path calls
    the device kernel through kc_run_split.)"""
    n, W = keys.shape
    off = [0]
    for s in splitters:
        lo, hi = 0, n
        while lo < hi:
            mid = (lo + hi) // 2
            if tuple(keys[mid]) < tuple(s):
                lo = mid + 1
            else:
                hi = mid
        off.append(lo)
    off.append(n)
    return np.asarray(off, dtype=np.int64)


# ------------------------------------------------------------------- exchange
def exchange_slices(keys_t, counts_t, offsets, group=None):
    """All-to-all-v of run slices.  keys_t [n, W] int64 and counts_t [n] int32 are this
    rank's run (any device); slice p = rows offsets[p]:offsets[p+1] goes to rank p.
    Returns (recv_keys, recv_counts, recv_sizes): the slices received, concatenated in
    source-rank order, and their lengths."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    send = [int(offsets[p + 1] - offsets[p]) for p in range(world)]
    dev = keys_t.device
    send_t = torch.tensor(send, dtype=torch.int64, device=dev)
    recv_t = torch.empty_like(send_t)
    backend = dist.get_backend(group)
    if backend == "nccl":
        dist.all_to_all_single(recv_t, send_t, group=group)
    else:
        _p2p_all_to_all(recv_t, send_t, [1] * world, [1] * world, group)
    recv = [int(x) for x in recv_t.tolist()]
    W = keys_t.shape[1]
    rk = torch.empty((sum(recv), W), dtype=keys_t.dtype, device=dev)
    rc = torch.empty((sum(recv),), dtype=counts_t.dtype, device=dev)
    if backend == "nccl":
        dist.all_to_all_single(rk, keys_t, output_split_sizes=recv, input_split_sizes=send, group=group)
        dist.all_to_all_single(rc, counts_t, output_split_sizes=recv, input_split_sizes=send, group=group)
    else:
        _p2p_all_to_all(rk, keys_t, recv, send, group)
        _p2p_all_to_all(rc, counts_t, recv, send, group)
    return rk, rc, recv


def _p2p_all_to_all(out, inp, recv_sizes, send_sizes, group):
    """all_to_all_single for backends without it (gloo on CPU): batched isend/irecv."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    so = np.concatenate([[0], np.cumsum(send_sizes)])
    ro = np.concatenate([[0], np.cumsum(recv_sizes)])
    out[ro[rank]:ro[rank + 1]] = inp[so[rank]:so[rank + 1]]
    ops = []
    for p in range(world):
        if p == rank:
            continue
        if send_sizes[p]:
            ops.append(dist.P2POp(dist.isend, inp[so[p]:so[p + 1]].contiguous(), dist.get_global_rank(group, p) if group else p, group))
        if recv_sizes[p]:
            ops.append(dist.P2POp(dist.irecv, out[ro[p]:ro[p + 1]], dist.get_global_rank(group, p) if group else p, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


# ---------------------------------------------------------------- device glue
class _CudaView:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (ptr, False), "shape": shape, "typestr": typestr, "version": 2}


def run_as_tensors(run, device):
    """Zero-copy torch views of a device-resident run: keys [n, W] int64, counts [n] int32."""
    import torch
    kptr, cptr, n = run.device_arrays()
    W = run._c.words
    if n == 0:
        return (torch.empty((0, W), dtype=torch.int64, device=device),
                torch.empty((0,), dtype=torch.int32, device=device))
    keys = torch.as_tensor(_CudaView(kptr, (n, W), "<i8"), device=device)
    counts = torch.as_tensor(_CudaView(cptr, (n,), "<i4"), device=device)
    return keys, counts


def count_shard(counter, d_reads_ptr, n_bytes, device, group=None, splitters=None, peer=None):
    """This rank's share of the distributed count.  Returns the Run holding the final
    records of this rank's key range.  Collective: every rank of `group` must call it.

    Runs of the partitioned hash path carry their partition structure (n_sub equal key ranges
    with record offsets); when all ranks agree on it, rank r owns ranges [r*n_sub/P, (r+1)*n_sub/P),
    receives each peer's records and offsets for them, and combines the P parts range by range in
    shared-memory tables (kc_merge_parts): one pass whose cost does not grow with P.  Otherwise
    (sort path, k > 32) slices are cut at key splitters and merged by the merge-path tree."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if peer is not None and world > 1:
        peer.place()                                # count straight into the staging buffer the peers read
    local = counter.count_device(d_reads_ptr, n_bytes)
    if world == 1:
        return local
    if peer is not None:
        merged = peer.combine(local)                # the exchange happens inside the combine kernel
        if merged is not None:
            return merged
    return exchange_and_combine(counter, local, device, group, splitters)


def exchange_and_combine(counter, local, device, group=None, splitters=None):
    """Second half of count_shard: `local` is this rank's run; returns the run of this rank's
    key range. Frees `local`."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    off_ptr, n_sub, pbits = local.parts()
    plan = torch.tensor([n_sub, pbits], dtype=torch.int64, device=device)
    plans = [torch.empty_like(plan) for _ in range(world)]
    dist.all_gather(plans, plan, group=group)
    agreed = counter.words == 1 and n_sub >= world and n_sub % world == 0 and all(bool((q == plan).all()) for q in plans)
    keys_t, counts_t = run_as_tensors(local, device)
    if agreed:
        per = n_sub // world
        off_t = torch.as_tensor(_CudaView(off_ptr, (n_sub + 1,), "<i4"), device=device)
        # offsets of every owner's ranges, relative to the start of its slice; the last entry of a
        # row is the slice length, so this one small all-to-all also tells the receiver the sizes
        rel = (off_t.unfold(0, per + 1, per) - off_t[0:n_sub:per].unsqueeze(1)).contiguous()
        recv_rel = torch.empty_like(rel)
        nccl = dist.get_backend(group) == "nccl"
        if nccl:
            dist.all_to_all_single(recv_rel, rel, group=group)
        else:
            _p2p_all_to_all(recv_rel, rel, [1] * world, [1] * world, group)
        send = rel[:, per].tolist()
        sizes = recv_rel[:, per].tolist()
        starts = off_t[0:n_sub:per].tolist()
        rk = torch.empty((sum(sizes), 1), dtype=keys_t.dtype, device=device)
        rc = torch.empty((sum(sizes),), dtype=counts_t.dtype, device=device)
        if nccl:
            # keys and counts of all peers in ONE grouped launch
            rank = dist.get_rank(group)
            ops, pos = [], 0
            for p in range(world):
                peer = dist.get_global_rank(group, p) if group else p
                a, b = starts[p], starts[p] + send[p]
                if p == rank:
                    rk[pos:pos + sizes[p]].copy_(keys_t[a:b])
                    rc[pos:pos + sizes[p]].copy_(counts_t[a:b])
                else:
                    if send[p]:
                        ops.append(dist.P2POp(dist.isend, keys_t[a:b], peer, group))
                        ops.append(dist.P2POp(dist.isend, counts_t[a:b], peer, group))
                    if sizes[p]:
                        ops.append(dist.P2POp(dist.irecv, rk[pos:pos + sizes[p]], peer, group))
                        ops.append(dist.P2POp(dist.irecv, rc[pos:pos + sizes[p]], peer, group))
                pos += sizes[p]
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
        else:
            bounds = [starts[p] for p in range(world)] + [starts[-1] + send[-1]]
            _p2p_all_to_all(rk, keys_t, sizes, [bounds[p + 1] - bounds[p] for p in range(world)], group)
            _p2p_all_to_all(rc, counts_t, sizes, [bounds[p + 1] - bounds[p] for p in range(world)], group)
        torch.cuda.current_stream().synchronize()
        local.free()
        kp, cp, op, pos = [], [], [], 0
        for src, sz in enumerate(sizes):
            kp.append(rk.data_ptr() + pos * 8 if sz else 0)
            cp.append(rc.data_ptr() + pos * 4 if sz else 0)
            op.append(recv_rel[src].data_ptr())
            pos += sz
        return counter.merge_parts(kp, cp, op, sizes, per, pbits)
    if splitters is None:
        splitters = range_splitters(world, counter.words)
    off = local.split(splitters)
    # the run was produced on the counter's stream and is complete (count_device synchronises)
    rk, rc, sizes = exchange_slices(keys_t, counts_t, off, group)
    torch.cuda.current_stream().synchronize()
    local.free()
    parts, pos = [], 0
    for sz in sizes:
        parts.append(counter.run_from_device(rk.data_ptr() + pos * counter.words * 8, rc.data_ptr() + pos * 4, sz))
        pos += sz
    merged = counter.merge(parts)
    for p in parts:
        p.free()
    return merged


# ------------------------------------------------ exchange fused into the combine kernel (peer memory)
class PeerCombine:
    """The all-to-all fused into the combine kernel over NVLink peer memory.

    Every rank keeps its run (keys, counts, range offsets) in a staging buffer whose CUDA IPC
    handle its peers have opened once.  A step is then: count locally -> copy the run into the
    staging buffer -> barrier -> kc_merge_parts reads, for every key range this rank owns, the
    P parts straight out of the peers' staging buffers (P2P loads through NVSwitch) while it
    combines them in shared memory -> barrier.  No exchange kernel, no receive buffer: the
    transfer happens inside the kernel that consumes it.  Needs partition-structured runs
    (64-bit keys); exchange_and_combine() over NCCL is the general path."""

    def __init__(self, counter, device, max_records, n_sub_max=1 << 20, group=None):
        import torch
        import torch.distributed as dist
        self.c, self.dev, self.group = counter, device, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.cap, self.n_sub_max = int(max_records), int(n_sub_max)
        self.loc_offs = None
        self._readers_pending = False
        # one allocation: keys u64[cap] | counts u32[cap] | range offsets u32[n_sub_max + 1]
        self._o_counts = 8 * self.cap
        self._o_offs = (12 * self.cap + 255) // 256 * 256
        self.base, handle = counter.peer_alloc(self._o_offs + 4 * (self.n_sub_max + 1))
        self.keys = torch.as_tensor(_CudaView(self.base, (self.cap,), "<i8"), device=device)
        self.counts = torch.as_tensor(_CudaView(self.base + self._o_counts, (self.cap,), "<i4"), device=device)
        self.offs = torch.as_tensor(_CudaView(self.base + self._o_offs, (self.n_sub_max + 1,), "<i4"), device=device)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, handle, group=group)
        # per rank: base pointer of its staging buffer as mapped into this process
        self.peer_base, err = [], None
        for r, h in enumerate(everyone):
            try:
                self.peer_base.append(self.base if r == self.rank else counter.peer_open(h))
            except Exception as e:                      # no peer access to that GPU
                err = e
                break
        # all ranks or none: a rank that cannot map a peer makes everybody fall back together
        ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            for r, p in enumerate(self.peer_base):
                if r != self.rank:
                    counter.peer_close(p)
            dist.barrier(group=group)
            self.keys = self.counts = self.offs = None
            counter.peer_free(self.base)
            self.base = None
            raise RuntimeError("peer staging buffers cannot be mapped on every rank: %s" % (err or "a peer failed"))

    def close(self):
        import torch.distributed as dist
        if self.base is None:
            return
        dist.barrier(group=self.group)                  # nobody is still reading
        for r, p in enumerate(self.peer_base):
            if r != self.rank:
                self.c.peer_close(p)
        dist.barrier(group=self.group)                  # every mapping is gone before the memory is
        self.keys = self.counts = self.offs = None
        self.c.peer_free(self.base)
        self.base = None

    def _settle(self):
        """Peers may still be reading this rank's staging buffer for the previous step: wait for them
        before it is written again. Deferred to here so that it overlaps the local count."""
        import torch.distributed as dist
        if self._readers_pending:
            dist.barrier(group=self.group)
            self._readers_pending = False

    def place(self):
        """Have the next run of the partitioned path written straight into the staging buffer
        (kc_place_next_run): saves the copy. Call right before the count whose run goes to combine()."""
        self._settle()
        self.c.place_next_run(self.base, self.base + self._o_counts, self.base + self._o_offs, self.cap, self.n_sub_max + 1)

    def combine(self, local):
        """`local`: this rank's run (freed here). Returns the run of this rank's key range, or None
        if the runs have no common partition structure (caller falls back to NCCL)."""
        import torch
        import torch.distributed as dist
        off_ptr, n_sub, pbits = local.parts()
        kptr, cptr, n = local.device_arrays()
        P = self.world
        usable = self.c.words == 1 and n_sub >= P and n_sub % P == 0 and n_sub <= self.n_sub_max and n <= self.cap
        per = n_sub // P if usable else 1
        placed = usable and kptr == self.base               # the count already wrote it here (place())
        if usable and not placed:
            self._settle()
            keys_t, counts_t = run_as_tensors(local, self.dev)
            self.keys[:n].copy_(keys_t[:, 0])
            self.counts[:n].copy_(counts_t)
            self.offs[:n_sub + 1].copy_(torch.as_tensor(_CudaView(off_ptr, (n_sub + 1,), "<i4"), device=self.dev))
        # One small all-gather: plan, size, and the record boundaries of the owners' ranges. It is
        # queued behind the copies above on this stream, so a peer that has received this rank's
        # entry knows that its staging buffer is complete: no separate barrier.
        info = torch.zeros(4 + P + 1, dtype=torch.int64, device=self.dev)
        info[0], info[1], info[2], info[3] = n_sub, pbits, n, int(usable)
        if usable:
            info[4:] = self.offs[:n_sub + 1:per].to(torch.int64)
        infos = [torch.empty_like(info) for _ in range(P)]
        dist.all_gather(infos, info, group=self.group)
        infos = [t.tolist() for t in infos]
        if not all(i[3] == 1 and i[0] == n_sub and i[1] == pbits for i in infos):
            return None                                 # (a placed run is still a valid run for the NCCL exchange)
        local.free()
        kp = list(self.peer_base)                       # absolute offsets index the peer's whole array
        cp = [b + self._o_counts for b in self.peer_base]
        # The range offsets are read twice per range and part: bring this rank's slice of every
        # peer's offsets over once (per + 1 integers each) instead of paying a link round trip per read.
        if self.loc_offs is None or self.loc_offs.shape[1] < per + 1:
            self.loc_offs = torch.empty((P, per + 1), dtype=torch.int32, device=self.dev)
        op = []
        for r, b in enumerate(self.peer_base):
            src = b + self._o_offs + 4 * self.rank * per
            if r == self.rank:
                op.append(src)
            else:
                self.loc_offs[r, :per + 1].copy_(torch.as_tensor(_CudaView(src, (per + 1,), "<i4"), device=self.dev))
                op.append(self.loc_offs[r].data_ptr())
        torch.cuda.current_stream().synchronize()
        sizes = [i[4 + self.rank + 1] - i[4 + self.rank] for i in infos]    # records this rank reads from each peer
        merged = self.c.merge_parts(kp, cp, op, sizes, per, pbits)
        self._readers_pending = True                    # settled (barrier) before the staging buffer is written again
        return merged

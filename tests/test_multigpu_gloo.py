"""World-size-2 test of the multi-GPU host logic on CPU (gloo): the all-gathered histogram, the
owner ranges cut from it (multigpu.plan_owner_ranges, the host model of the device plan) and the
all-to-all-v exchange, with the oracle standing in for the device kernels.  The device path
(kc_xchg_*: local count, pull over peer memory, fold) is covered by the gpu-marked tests below:
all ranks as contexts of one process on one GPU (runs on the driver's one-GPU box), and one
process per GPU over NCCL when the box has several."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, k, L, R, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle
    from kmer_counter_b200 import multigpu
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W, S = oracle.words(k), oracle.record_size(k)
    # this rank's shard of the reads, counted locally (oracle = stand-in for kc_count_device)
    per = R // world
    reads = oracle.gen_reads(per, L, 30000, 0.01, 0.003, seed=5, first_read=rank * per)
    local = oracle.process_chunk(reads, L, k)
    keys, counts = oracle.records_to_arrays(local, k)
    # 1024-bin histogram of the records' leading 10 bits, all-gathered; every rank derives the same
    # owner ranges from it (what x_plan_kernel does on the device)
    hist = np.bincount((keys[:, 0] >> np.uint64(54)).astype(np.int64), minlength=multigpu.XCHG_BINS).astype(np.int32)
    hist_t = torch.from_numpy(hist)
    all_t = [torch.empty_like(hist_t) for _ in range(world)]
    dist.all_gather(all_t, hist_t)
    lo = multigpu.plan_owner_ranges(np.stack([t.numpy() for t in all_t]), world)
    assert lo[0] == 0 and lo[world] == multigpu.XCHG_BINS and all(lo[i] <= lo[i + 1] for i in range(world))
    sp = np.zeros((world - 1, W), dtype=np.uint64)
    for o in range(1, world):
        sp[o - 1, 0] = np.uint64(lo[o]) << np.uint64(54) if lo[o] < multigpu.XCHG_BINS else np.uint64(2**64 - 1)
    off = multigpu.slice_offsets_host(keys, sp)
    keys_t = torch.from_numpy(keys.view(np.int64).copy())
    counts_t = torch.from_numpy(counts.view(np.int32).copy())
    rk, rc, sizes = multigpu.exchange_slices(keys_t, counts_t, off)
    # merge what arrived (oracle = stand-in for kc_merge_runs)
    runs, pos = [], 0
    for sz in sizes:
        kk = rk[pos:pos + sz].numpy().view(np.uint64).reshape(-1, W)
        cc = rc[pos:pos + sz].numpy().view(np.uint32)
        rec = np.zeros((sz, S), dtype=np.uint8)
        rec[:, :8 * W] = kk.view(np.uint8).reshape(sz, 8 * W)
        rec[:, 8 * W:] = cc.view(np.uint8).reshape(sz, 4)
        runs.append(rec.tobytes())
        pos += sz
    mine = oracle.merge_runs(runs, k)
    mk, _ = oracle.records_to_arrays(mine, k)
    if len(mk):
        b = (mk[:, 0] >> np.uint64(54)).astype(np.int64)
        assert ((b >= lo[rank]) & (b < lo[rank + 1])).all()
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        q.put(b"".join(gathered))
    dist.destroy_process_group()


@pytest.mark.parametrize("k,L", [(31, 100), (63, 100)])
def test_two_rank_exchange_matches_single_process(k, L):
    import torch.multiprocessing as mp
    import oracle
    R, world = 2000, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, k, L, R, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    reads = oracle.gen_reads(R, L, 30000, 0.01, 0.003, seed=5)
    assert got == oracle.count(reads, L, k)


def _gpu_worker(rank, world, port, k, L, R, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import kmer_counter_b200 as kc
    from kmer_counter_b200 import multigpu, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    per = R // world
    d = torch.empty(per * L + 256, dtype=torch.uint8, device=dev)
    synth.synth_reads_device(d.data_ptr(), per, L, 30000, 0.01, 0.003, seed=5, first_read=rank * per)
    torch.cuda.synchronize()
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    with kc.Counter(k, L, device=rank, stream=stream.cuda_stream) as c:
        ex = multigpu.Exchange(c, dev, per)
        mine = None
        for _ in range(2):                             # twice: the workspaces are reused
            run = multigpu.count_shard(c, d.data_ptr(), per * L, dev, exchange=ex)
            got = run.to_bytes()
            run.free()
            assert mine is None or got == mine
            mine = got
        # the same through the NCCL all-to-all-v of run slices (the path for k > 64)
        run = multigpu.count_shard(c, d.data_ptr(), per * L, dev)
        nccl = run.to_bytes()
        run.free()
    gathered, gathered2 = [None] * world, [None] * world
    dist.all_gather_object(gathered, mine)
    dist.all_gather_object(gathered2, nccl)
    if rank == 0:
        q.put((b"".join(gathered), b"".join(gathered2)))
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("k", [31, 63])
def test_multi_gpu_count_matches_oracle(k):
    """One process per GPU (gpurun --gpus 2/4/8); uses all of them, up to 8. Skipped on a one-GPU box,
    where test_exchange_all_ranks_on_one_device covers the same kernels."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import oracle
    world = min(8, 1 << (torch.cuda.device_count().bit_length() - 1))
    R, L = 8000, 100
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_worker, args=(r, world, port, k, L, R, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, got_nccl = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    reads = oracle.gen_reads(R, L, 30000, 0.01, 0.003, seed=5)
    want = oracle.count(reads, L, k)
    assert got == want
    assert got_nccl == want


@pytest.mark.gpu
@pytest.mark.parametrize("P,R,L,k,G,e,n,sub_target", [
    (2, 20000, 100, 31, 100000, 0.01, 0.002, 0),
    (4, 30000, 100, 31, 50000, 0.01, 0.001, 0),
    (8, 40000, 100, 31, 0, 0.0, 0.0, 0),          # iid reads: every record distinct, nothing to fold
    (3, 9000, 100, 63, 40000, 0.001, 0.001, 0),   # 128-bit keys, a rank count that is no power of two
    (8, 800, 100, 31, 3000, 0.0, 0.01, 0),        # tiny shards: ranks with few or no buckets
    (4, 12000, 70, 28, 30000, 0.01, 0.001, 0),    # masked tail
    (5, 7, 100, 31, 0, 0.0, 0.0, 0),              # fewer reads than ranks
    (2, 40000, 100, 31, 0, 0.0, 0.0, 1),       # 11-bit level 2 (a job of > 2^20 x 1433 records): 2^20 sub-buckets per rank
    (4, 80000, 100, 31, 0, 0.0, 0.0, 1),       # ... whose sub-buckets exceed the target: S3c's large variant, chosen on the device
    (3, 30000, 100, 63, 0, 0.0, 0.001, 1),     # ... 128-bit keys
])
def test_exchange_all_ranks_on_one_device(monkeypatch, P, R, L, k, G, e, n, sub_target):
    """The multi-GPU exchange with all P ranks as contexts of this process on device 0
    (kc_xchg_run_all: nothing waits inside a kernel, ordering is by events). Rank r must end with
    the r-th key range, the concatenation must be the oracle's artefact."""
    import torch
    import oracle
    import kmer_counter_b200 as kc
    from kmer_counter_b200 import engine
    reads = oracle.gen_reads(R, L, G, e, n, seed=P * 1000 + k)
    want = oracle.count(reads, L, k)
    per = (R + P - 1) // P
    if sub_target:
        monkeypatch.setenv("KC_SW_SUB_TARGET", str(sub_target))
    cs = [kc.Counter(k, L, method="super") for _ in range(P)]
    try:
        bufs = []
        for r, c in enumerate(cs):
            c.xchg_begin(r, P, max(per, 16))
            part = reads[r * per * L:(r + 1) * per * L]
            d = torch.from_numpy(part.copy()).cuda() if len(part) else torch.empty(16, dtype=torch.uint8, device="cuda")
            bufs.append(d)
            c.accum_add_device(d.data_ptr(), len(part))
        for rep in range(2):                           # the second round counts nothing: empty runs
            runs = engine.xchg_run_all(cs)
            parts = [r.to_bytes() for r in runs]
            info = [c.xchg_info(P) for c in cs]
            for r in runs:
                r.free()
            if rep == 1:
                assert all(len(x) == 0 for x in parts)
                break
            assert b"".join(parts) == want
            lo = info[0][0]
            assert all(i[0] == lo for i in info) and lo[0] == 0 and lo[P] == 1024
            S = 8 * ((k + 31) // 32) + 4
            for r, x in enumerate(parts):              # rank r holds exactly the keys of buckets [lo[r], lo[r+1])
                if x:
                    keys, _ = oracle.records_to_arrays(x, k)
                    b = (keys[:, 0] >> np.uint64(54)).astype(np.int64)
                    assert b.min() >= lo[r] and b.max() < lo[r + 1]
            assert sum(i[1] for i in info) >= len(want) // S
    finally:
        for c in cs:
            c.close()

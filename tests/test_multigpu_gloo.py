"""World-size-2 test of the multi-GPU host logic on CPU (gloo): ownership, slice offsets and
the all-to-all-v exchange, with the oracle standing in for the device kernels.  The real
device path (NCCL, kc_run_split, merge-path) is covered by the gpu-marked test below."""
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
    sp = multigpu.range_splitters(world, W)
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
        assert (multigpu.owner_of(mk[:, 0], world) == rank).all()
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
    with kc.Counter(k, L, device=rank) as c:
        run = multigpu.count_shard(c, d.data_ptr(), per * L, dev)
        mine = run.to_bytes()
        run.free()
        if k <= 32:                                    # the same through the peer-memory fused combine
            peer = multigpu.PeerCombine(c, dev, max_records=per * (L - k + 1) + 16, n_sub_max=1 << 16)
            for _ in range(2):                         # twice: the staging buffers are reused
                run = multigpu.count_shard(c, d.data_ptr(), per * L, dev, peer=peer)
                assert run.to_bytes() == mine, "peer-memory combine differs from the NCCL path"
                run.free()
            peer.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        q.put(b"".join(gathered))
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("k", [31, 63])
def test_two_gpu_count_matches_oracle(k):
    """Needs at least two devices (gpurun --gpus 2/4/8); uses all of them, up to 8. Skipped on a one-GPU box."""
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
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    reads = oracle.gen_reads(R, L, 30000, 0.01, 0.003, seed=5)
    assert got == oracle.count(reads, L, k)

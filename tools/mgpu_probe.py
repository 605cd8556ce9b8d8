"""Phase timings of the multi-GPU step (development aid). torchrun --nproc-per-node N tools/mgpu_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import kmer_counter_b200 as kc
from kmer_counter_b200 import multigpu, synth

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
R, L, k = 10_000_000, 100, 31
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
c = kc.Counter(k, L, device=lr, stream=stream.cuda_stream)
d = torch.empty(R * L + 256, dtype=torch.uint8, device=dev)
synth.synth_reads_device(d.data_ptr(), R, L, 100_000_000, 1e-3, 0.0, 2, first_read=rank * R, stream=stream.cuda_stream)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(4):
    dist.barrier(); t0 = T()
    local = c.count_device(d.data_ptr(), R * L); t1 = T()
    off_ptr, n_sub, pbits = local.parts()
    keys_t, counts_t = multigpu.run_as_tensors(local, dev)
    per = n_sub // world
    off_t = torch.as_tensor(multigpu._CudaView(off_ptr, (n_sub + 1,), "<i4"), device=dev)
    bounds = off_t[::per].to(torch.int64)
    rel = torch.stack([off_t[p * per:(p + 1) * per + 1] - off_t[p * per] for p in range(world)]).contiguous()
    recv_rel = torch.empty_like(rel); dist.all_to_all_single(recv_rel, rel); t2 = T()
    rk, rc, sizes = multigpu.exchange_slices(keys_t, counts_t, bounds.tolist()); t3 = T()
    local.free()
    kp, cp, op, pos = [], [], [], 0
    for src, sz in enumerate(sizes):
        kp.append(rk.data_ptr() + pos * 8); cp.append(rc.data_ptr() + pos * 4); op.append(recv_rel[src].data_ptr()); pos += sz
    m = c.merge_parts(kp, cp, op, sizes, per, pbits); t4 = T()
    n = len(m); m.free(); t5 = T()
    if rank == 0:
        print("it %d: count %.2f  offsets+a2a %.2f  exchange %.2f (%.0f MB sent)  merge_parts %.2f  free %.2f  total %.2f ms  records %d" %
              (it, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, keys_t.numel()*12/1e6*(world-1)/world, (t4-t3)*1e3, (t5-t4)*1e3, (t5-t0)*1e3, n), flush=True)
dist.destroy_process_group()

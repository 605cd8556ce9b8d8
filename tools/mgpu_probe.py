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
    t2 = t3 = t1
    m = multigpu.exchange_and_combine(c, local, dev); t4 = T()
    n = len(m); m.free(); t5 = T()
    if rank == 0:
        print("it %d: count %.2f  offsets+a2a %.2f  exchange %.2f (%.0f MB sent)  merge_parts %.2f  free %.2f  total %.2f ms  records %d" %
              (it, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, 0, (t4-t3)*1e3, (t5-t4)*1e3, (t5-t0)*1e3, n), flush=True)
dist.destroy_process_group()

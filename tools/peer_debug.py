"""Step-by-step check of the peer-memory combine (development aid). torchrun --nproc-per-node 2 tools/peer_debug.py"""
import os, sys, time, faulthandler
faulthandler.dump_traceback_later(40, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import kmer_counter_b200 as kc
from kmer_counter_b200 import multigpu, synth
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
def say(*a):
    print("[rank %d]" % rank, *a, flush=True)
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
R, L, k = 200_000, 100, 31
c = kc.Counter(k, L, device=lr)
d = torch.empty(R * L + 256, dtype=torch.uint8, device=dev)
synth.synth_reads_device(d.data_ptr(), R, L, 1_000_000, 1e-3, 0.0, 2, first_read=rank * R)
torch.cuda.synchronize(); say("generated")
ref = multigpu.count_shard(c, d.data_ptr(), R * L, dev); want = ref.to_bytes(); ref.free(); say("nccl path ok", len(want))
peer = multigpu.PeerCombine(c, dev, max_records=R * 70 + 16, n_sub_max=1 << 16); say("peer buffers mapped", [hex(b) for b in peer.peer_base])
import ctypes as C
lib = kc._lib.load()
lib.kc_debug_probe_read.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_ulonglong)]   # not part of kc_api.h
peer.keys[:1000] = torch.arange(1000, device=dev) + 1000 * rank
torch.cuda.synchronize(); dist.barrier()
for r, pb in enumerate(peer.peer_base):
    out = C.c_ulonglong()
    rc = lib.kc_debug_probe_read(pb, 1000, C.byref(out))
    say("kernel read of rank %d's staging (ptr %x): rc=%d sum=%d expect=%d" % (r, pb, rc, out.value, 499500 + 1000 * 1000 * r))
dist.barrier()
for it in range(2):
    run = multigpu.count_shard(c, d.data_ptr(), R * L, dev, peer=peer)
    got = run.to_bytes(); run.free()
    say("peer path iteration", it, "equal:", got == want)
dist.barrier(); dist.destroy_process_group(); say("done")

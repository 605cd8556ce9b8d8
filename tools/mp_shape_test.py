"""The P-GPU combine's shape on one GPU (development aid): P shards counted locally, then rank 0's share
(the first n_sub/P key ranges of every run) combined with kc_merge_parts from local memory. Comparing its
time with the multi-GPU step separates the kernel's own cost from the NVLink reads."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kmer_counter_b200 as kc
from kmer_counter_b200 import synth, multigpu

P = int(os.environ.get("P", "8"))
R, L, k = 10_000_000, 100, 31
dev = torch.device("cuda", 0)
d = torch.empty(R * L + 256, dtype=torch.uint8, device=dev)
with kc.Counter(k, L) as c:
    runs = []
    for r in range(P):
        synth.synth_reads_device(d.data_ptr(), R, L, 100_000_000, 1e-3, 0.0, 2, first_read=r * R)
        torch.cuda.synchronize()
        runs.append(c.count_device(d.data_ptr(), R * L))
    off_ptr, n_sub, pbits = runs[0].parts()
    per = n_sub // P
    kp, cp, op, sizes = [], [], [], []
    for run in runs:
        o, ns, pb = run.parts()
        assert ns == n_sub and pb == pbits
        kptr, cptr, n = run.device_arrays()
        off = torch.as_tensor(multigpu._CudaView(o, (n_sub + 1,), "<i4"), device=dev)
        kp.append(kptr); cp.append(cptr); op.append(o); sizes.append(int(off[per]) - int(off[0]))
    for it in range(3):
        torch.cuda.synchronize(); t = time.perf_counter()
        m = c.merge_parts(kp, cp, op, sizes, per, pbits)
        torch.cuda.synchronize()
        print("P=%d: %d records in -> %d out, %.2f ms" % (P, sum(sizes), len(m), (time.perf_counter() - t) * 1e3))
        m.free()

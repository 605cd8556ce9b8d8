"""Throughput of the run merges (development aid): 4 shards of C2 shape counted on one GPU, then
kc_merge_runs over 2 (merge path) and 4 of them (same-plan fast path), and kc_merge_parts directly."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kmer_counter_b200 as kc
from kmer_counter_b200 import synth

R, L, k = 10_000_000, 100, 31
dev = torch.device("cuda", 0)
d = torch.empty(R * L + 256, dtype=torch.uint8, device=dev)
with kc.Counter(k, L) as c:
    runs = []
    for r in range(4):
        synth.synth_reads_device(d.data_ptr(), R, L, 100_000_000, 1e-3, 0.0, 2, first_read=r * R)
        torch.cuda.synchronize()
        runs.append(c.count_device(d.data_ptr(), R * L))
    for n in (2, 4):
        for it in range(2):
            torch.cuda.synchronize(); t = time.perf_counter()
            m = c.merge(runs[:n])
            torch.cuda.synchronize(); dt = time.perf_counter() - t
            n_in = sum(len(r) for r in runs[:n])
            print("kc_merge_runs,   %d runs: %d records in -> %d out, %.2f ms, %.1f G records/s" % (n, n_in, len(m), dt * 1e3, n_in / dt / 1e9))
            m.free()
        off, n_sub, pbits = runs[0].parts()
        kp, cp, op, sz = [], [], [], []
        for r in runs[:n]:
            o, ns, pb = r.parts(); a, b, cnt = r.device_arrays()
            kp.append(a); cp.append(b); op.append(o); sz.append(cnt)
        for it in range(2):
            torch.cuda.synchronize(); t = time.perf_counter()
            m = c.merge_parts(kp, cp, op, sz, n_sub, pbits)
            torch.cuda.synchronize(); dt = time.perf_counter() - t
            print("merge_parts,     %d runs: %d records in -> %d out, %.2f ms, %.1f G records/s" % (n, sum(sz), len(m), dt * 1e3, sum(sz) / dt / 1e9))
            m.free()

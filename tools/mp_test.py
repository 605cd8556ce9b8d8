import sys, os, time
sys.path.insert(0, '/root/repo')
import torch, numpy as np
import kmer_counter_b200 as kc
from kmer_counter_b200 import synth, multigpu
R, L, k = 10_000_000, 100, 31
dev = torch.device('cuda', 0)
d = torch.empty(R * L + 256, dtype=torch.uint8, device=dev)
synth.synth_reads_device(d.data_ptr(), R, L, 100_000_000, 1e-3, 0.0, 2)
torch.cuda.synchronize()
with kc.Counter(k, L) as c:
    run = c.count_device(d.data_ptr(), R * L)
    off_ptr, n_sub, pbits = run.parts()
    kp, cp, n = run.device_arrays()
    for nsrc in (1, 2, 1):
        for it in range(2):
            torch.cuda.synchronize(); t = time.perf_counter()
            m = c.merge_parts([kp] * nsrc, [cp] * nsrc, [off_ptr] * nsrc, [n] * nsrc, n_sub, pbits)
            torch.cuda.synchronize(); print(nsrc, 'sources: %.2f ms' % ((time.perf_counter() - t) * 1e3), len(m), n)
            if nsrc == 2 and it == 0:
                k1, c1 = multigpu.run_as_tensors(run, dev); k2, c2 = multigpu.run_as_tensors(m, dev)
                print('keys equal', bool((k1 == k2).all()), 'counts doubled', bool((c2 == 2 * c1).all()))
            m.free()

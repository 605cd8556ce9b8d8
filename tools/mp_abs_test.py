import sys, os
sys.path.insert(0, '/root/repo')
import torch, numpy as np
import kmer_counter_b200 as kc
from kmer_counter_b200 import synth, multigpu
R, L, k = 200_000, 100, 31
dev = torch.device('cuda', 0)
d = torch.empty(R * L + 256, dtype=torch.uint8, device=dev)
synth.synth_reads_device(d.data_ptr(), R, L, 1_000_000, 1e-3, 0.0, 2)
torch.cuda.synchronize()
with kc.Counter(k, L) as c:
    run = c.count_device(d.data_ptr(), R * L)
    off_ptr, n_sub, pbits = run.parts()
    kp, cp, n = run.device_arrays()
    off = torch.as_tensor(multigpu._CudaView(off_ptr, (n_sub + 1,), "<i4"), device=dev)
    per = n_sub // 2
    outs = []
    for half in range(2):
        sizes = [int(off[(half + 1) * per] - off[half * per])] * 2
        m = c.merge_parts([kp, kp], [cp, cp], [off_ptr + 4 * half * per] * 2, sizes, per, pbits)   # absolute offsets, two copies
        outs.append(m)
        print('half', half, 'records', len(m), 'expected', sizes[0])
    k1, c1 = multigpu.run_as_tensors(run, dev)
    ka = torch.cat([multigpu.run_as_tensors(m, dev)[0] for m in outs]); ca = torch.cat([multigpu.run_as_tensors(m, dev)[1] for m in outs])
    print('keys equal', bool((ka == k1).all()), 'counts doubled', bool((ca == 2 * c1).all()))

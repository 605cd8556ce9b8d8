import sys, os, time
sys.path.insert(0, '/root/repo')
import torch, numpy as np
import kmer_counter_b200 as kc
import oracle
R, L, k = 5_000_000, 100, 31
fq = np.frombuffer(oracle.gen_fastq(R, L, 100_000_000, 1e-3, 0.0, seed=2), dtype=np.uint8)
dev = torch.device('cuda', 0)
d_text = torch.from_numpy(fq.copy()).cuda()
d_reads = torch.empty(R * L, dtype=torch.uint8, device=dev)
with kc.Counter(k, L, max_chunk_bytes=R * L, n_slots=2) as c:
    for it in range(3):
        torch.cuda.synchronize(); t = time.perf_counter()
        n, used, fl = c.parse_fastq_device(d_text.data_ptr(), d_text.numel(), d_reads.data_ptr(), d_reads.numel())
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        print('device parse: %d reads, %.2f ms, %.0f GB/s of FASTQ text' % (n, dt * 1e3, fq.size / dt / 1e9), fl)
    pin = c.host_alloc(fq.size); pin[:] = fq
    for it in range(3):
        t = time.perf_counter(); used, fl = c.submit_fastq(0, pin); run = c.wait(0); dt = time.perf_counter() - t
        print('host FASTQ text -> counted run: %.1f ms (%.2f G k-mers/s, %.1f GB/s of text), records %d' % (dt * 1e3, R * 70 / dt / 1e9, fq.size / dt / 1e9, len(run)))
        run.free()

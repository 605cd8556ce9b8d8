"""Single process driving P GPUs through kc_xchg_*: per-rank stage times of the exchange path.
usage: xchg_probe.py P [reads_per_gpu] [seq]   (seq: run the ranks' count_local one after the other)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import kmer_counter_b200 as kc  # noqa: E402
from kmer_counter_b200 import engine, synth  # noqa: E402

P = int(sys.argv[1])
R = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
seq = len(sys.argv) > 3 and sys.argv[3] == "seq"
L, k = 100, 31
ndev = torch.cuda.device_count()
cs, bufs = [], []
for r in range(P):
    d = r % ndev
    torch.cuda.set_device(d)
    buf = torch.empty(R * L + 256, dtype=torch.uint8, device="cuda:%d" % d)
    synth.synth_reads_device(buf.data_ptr(), R, L, 100_000_000 * P, 1e-3, 0.0, 2, first_read=r * R)
    torch.cuda.synchronize()
    bufs.append(buf)
    c = kc.Counter(k, L, device=d, method="super")
    c.xchg_begin(r, P, R)
    cs.append(c)
for it in range(3):
    t0 = time.perf_counter()
    for r, c in enumerate(cs):
        c.accum_add_device(bufs[r].data_ptr(), R * L)
    for c in cs:
        c.sync()
    t1 = time.perf_counter()
    if seq:
        for r, c in enumerate(cs):
            for q, o in enumerate(cs):
                if q != r:
                    c.xchg_set_peer(q, o)
        for c in cs:
            c.xchg_count_local()
            c.sync()
        print("count_local one rank at a time: %.2f ms each" % ((time.perf_counter() - t1) * 1e3 / P))
        sys.exit(0)
    runs = engine.xchg_run_all(cs)
    t2 = time.perf_counter()
    print("iter %d: scatter %.2f ms, exchange+count %.2f ms, records %s" % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, [len(x) for x in runs]))
    for r, c in enumerate(cs):
        st = c.stats()
        print("   rank %d" % r, dict(zip(st["stage_names"], [round(x, 3) for x in st["ms_stage"]])), c.xchg_info(P)[1:])
    for x in runs:
        x.free()

"""Stage timings of one device-resident chunk (development aid, not the bench)."""
import argparse, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import kmer_counter_b200 as kc, oracle

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=10_000_000)
ap.add_argument("--k", type=int, default=31)
ap.add_argument("--L", type=int, default=100)
ap.add_argument("--genome", type=int, default=100_000_000)
ap.add_argument("--method", default="sort")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--check", type=int, default=0)
ap.add_argument("--target", type=int, default=0)
a = ap.parse_args()
t = time.time()
reads = oracle.gen_reads(a.reads, a.L, a.genome, 1e-3, 0.0, seed=2)
print("gen %.1fs" % (time.time() - t), flush=True)
d = torch.from_numpy(reads).cuda()
with kc.Counter(a.k, a.L, method=a.method, table_slots=a.target) as c:
    for it in range(a.iters):
        torch.cuda.synchronize(); t = time.time()
        run = c.count_device(d.data_ptr(), d.numel())
        torch.cuda.synchronize(); dt = time.time() - t
        st = c.stats()
        n = a.reads * (a.L - a.k + 1)
        print("iter %d wall %.2f ms total %.2f U=%d -> %.1f Gkmer/s | " % (it, dt * 1e3, st["ms_total"], len(run), n / dt / 1e9)
              + "  ".join("%s %.2fms %.0fGB/s" % (nm, ms, b / ms / 1e6 if ms > 0 else 0)
                          for nm, ms, b in zip(st["stage_names"], st["ms_stage"], st["stage_bytes"])), flush=True)
        if it == 0 and a.check:
            sub = reads[: a.check * a.L]
            want = oracle.count(sub, a.L, a.k, threads=8)
            r2 = c.count_device(d.data_ptr(), sub.size)
            print("prefix parity:", r2.to_bytes() == want, flush=True)
            r2.free()
        run.free()

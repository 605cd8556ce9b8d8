"""Runs the counting path a few times on a synthetic device-resident input (ncu / timing driver).
usage: super_one.py [method] [reads] [k] [iters] [genome] [zipf_loci]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import kmer_counter_b200 as kc  # noqa: E402
from kmer_counter_b200 import synth  # noqa: E402

method = sys.argv[1] if len(sys.argv) > 1 else "super"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 31
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
G = int(sys.argv[5]) if len(sys.argv) > 5 else 100_000_000
zipf = int(sys.argv[6]) if len(sys.argv) > 6 else 0
L = 100
dev = torch.device("cuda", 0)
d_reads = torch.empty(R * L + 256, dtype=torch.uint8, device=dev)
synth.synth_reads_device(d_reads.data_ptr(), R, L, G, 1e-3, 0.0, 2, zipf_loci=zipf)
torch.cuda.synchronize()
with kc.Counter(k, L, method=method, table_slots=int(os.environ.get("KC_TABLE_SLOTS", "0"))) as c:
    for i in range(iters):
        run = c.count_device(d_reads.data_ptr(), R * L)
        n = len(run)
        run.free()
        st = c.stats()
        print(method, "records", n, "ms_total %.3f" % st["ms_total"],
              dict(zip(st["stage_names"], [round(x, 3) for x in st["ms_stage"]])), flush=True)
    if st["method_used"] == "super":
        print("scalars", c.debug_scalars(), flush=True)

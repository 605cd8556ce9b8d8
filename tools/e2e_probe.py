"""Timeline of the end-to-end pipeline (development aid): per-call wall times of submit / wait / copy."""
import os, sys, time, threading, queue
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kmer_counter_b200 as kc
from kmer_counter_b200 import synth

R, L, k = 10_000_000, 100, 31
NS = int(os.environ.get("NS", "3"))
steps = 12
dev = torch.device("cuda:0")
d = torch.empty(R * L, dtype=torch.uint8, device=dev)
synth.synth_reads_device(d.data_ptr(), R, L, 10_000_000, 0.01, 0.003, seed=1)
torch.cuda.synchronize()
host = d.cpu().numpy()
c = kc.Counter(k, L, device=0, n_slots=NS, max_chunk_bytes=R * L)
for sl in range(NS):
    c.slot_buffer(sl)[:R * L] = host
cap = 200_000_000 * 12
out = c.host_alloc(cap)
T0 = time.perf_counter()
log = []
def ev(name, t0):
    log.append((name, (t0 - T0) * 1e3, (time.perf_counter() - t0) * 1e3))
for _ in range(3):
    c.submit(0, R * L); r = c.wait(0); r.copy_into(out.ctypes.data, cap); r.free()
todo = queue.Queue(maxsize=1)
def reader():
    while True:
        run = todo.get()
        if run is None: return
        t = time.perf_counter(); run.copy_into(out.ctypes.data, cap); ev("copy", t)
        t = time.perf_counter(); run.free(); ev("free", t)
th = threading.Thread(target=reader); th.start()
torch.cuda.synchronize()
T0 = time.perf_counter()
depth = NS - 1
def take(sl):
    t = time.perf_counter(); run = c.wait(sl); ev("wait%d" % sl, t)
    t = time.perf_counter(); todo.put(run); ev("put", t)
for i in range(steps):
    t = time.perf_counter(); c.submit(i % NS, R * L); ev("submit%d" % (i % NS), t)
    if i >= depth: take((i - depth) % NS)
for i in range(max(steps - depth, 0), steps): take(i % NS)
todo.put(None); th.join()
torch.cuda.synchronize()
tot = (time.perf_counter() - T0) * 1e3
print("NS", NS, "total ms/step", tot / steps)
for name, at, dur in sorted(log, key=lambda x: x[1]):
    print("%8.1f  %-9s %7.2f" % (at, name, dur))

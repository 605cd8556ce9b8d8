#!/usr/bin/env python
"""bench.py -- k-mers counted per second at k=31 (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            (ours; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...   (the reference's CPU path)

A step is one pass of the whole counting path over one batch of synthetic reads of
BASELINE.json configs[1]'s shape (10M x 100 bp, k=31; per GPU when N > 1 -- weak
scaling): extract -> count -> sorted key-unique run [-> exchange -> merge].
  value : whole-job k-mers/s with the reads already resident in HBM
  e2e   : the same through the host-buffer API: pinned host reads -> H2D -> count ->
          packed records D2H into pinned host memory, every step
  roofline      : dominant kernel (radix scatter pass / hash insert), algorithmic bytes per
                  launch over its device time (CUDA events inside libkc_b200 on the stream
                  the kernels run on), against MEASURED_PEAKS.json's HBM copy bandwidth
  path_roofline : SURVEY.md 8(d)'s whole-path figure B_alg = B_in + N*(Kb+8) + U*S over step time
  cpu_baseline  : oracle/_ref (the reference's own sources) on a bounded sample, host cores
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "k-mers counted/s at k=31"
UNIT = "kmers/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU (configs[1]: 10M)")
    ap.add_argument("--read-len", type=int, default=100)
    ap.add_argument("--k", type=int, default=31)
    ap.add_argument("--genome", type=int, default=100_000_000)
    ap.add_argument("--sub-rate", type=float, default=1e-3)
    ap.add_argument("--n-rate", type=float, default=0.0)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--zipf-loci", type=int, default=0)
    ap.add_argument("--method", default="auto", choices=["auto", "sort", "hash"])
    ap.add_argument("--cpu-sample-reads", type=int, default=400_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--check", action="store_true", help="verify a prefix against the oracle before timing")
    ap.add_argument("--target", type=int, default=0, help="keys per sub-bucket the partition plan aims for (0 = library default)")
    return ap.parse_args()


E2E_SLOTS = 2           # pinned input slots of the end-to-end pipeline (chunks whose H2D / kernels are in flight)


def workload(a):
    return {"workload": "configs[1]: synthetic 10M x 100 bp reads, k=31, one B200" if a.gpus == 1 and a.reads == 10_000_000
            else "configs[1] shape per GPU (weak scaling)" if a.reads == 10_000_000 else "custom",
            "reads_per_gpu": a.reads, "read_len": a.read_len, "k": a.k, "genome_len": a.genome,
            "sub_rate": a.sub_rate, "n_rate": a.n_rate, "seed": a.seed, "zipf_loci": a.zipf_loci,
            "kmers_per_step_per_gpu": a.reads * (a.read_len - a.k + 1),
            "l2_policy": "inputs larger than L2 (1 GB of reads, 5.6 GB of keys per step vs 126 MB L2)"}


def bind_near_gpu(index):
    """Run this process on the CPUs next to GPU `index` (NVML's ideal affinity) before anything is
    allocated: the pinned buffers of the end-to-end path then live in that NUMA node's memory, which is
    what the GPU's PCIe link reaches without crossing the socket interconnect. Returns the CPU count
    bound to, or None when NVML has no answer (KC_NO_BIND=1 skips it)."""
    if os.environ.get("KC_NO_BIND") == "1":
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every 20 ms while the timed region runs."""

    def __init__(self, index):
        self.index, self.sm, self.mx, self.reasons, self.stop, self.th = index, [], 0, set(), threading.Event(), None
        self.nv = self.h = None
        try:                                        # NVML is initialised before the timed region starts
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            self.flags = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                          "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                          "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                          "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        except Exception as e:                      # no NVML: report that instead of inventing numbers
            self.reasons.add("nvml_unavailable: %s" % type(e).__name__)

    def _sample(self):
        if self.h is None:
            return
        self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
        r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in self.flags.items():
            if r & bit:
                self.reasons.add(name)

    def _run(self):
        while not self.stop.is_set():
            self._sample()
            self.stop.wait(0.005)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self._sample()                              # one sample is taken while the last step is still in flight
        self.stop.set()
        self.th.join(timeout=5)

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx or None,
                "reasons": sorted(self.reasons), "samples": len(sm)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# --------------------------------------------------------------- reference arm
def run_reference(a, rank):
    """The reference's own CPU implementation of the path (oracle/_ref: its sources compiled in
    place) on the host cores, on a bounded sample of this arm's workload."""
    if rank != 0:
        return
    import numpy as np
    import oracle
    cores = os.cpu_count() or 1
    L, k = a.read_len, a.k
    sample = min(a.reads, a.cpu_sample_reads)
    kind = "reference" if oracle.ref_available() else "port"
    reads = oracle.gen_reads(sample, L, a.genome, a.sub_rate, a.n_rate, seed=a.seed, zipf_loci=a.zipf_loci)
    tmp_root = "/dev/shm" if os.path.isdir("/dev/shm") else None

    def one_step():
        t0 = time.perf_counter()
        if kind == "reference":
            with tempfile.TemporaryDirectory(dir=tmp_root) as d:
                oracle.ref_count_packed(reads, L, k, 89364, cores, d, os.path.join(d, "out.bin"))
        else:
            oracle.count(reads, L, k, 89364, threads=cores)
        return time.perf_counter() - t0

    for _ in range(min(a.warmup, 1)):
        one_step()
    times = [one_step() for _ in range(max(a.steps, 1))]
    dt = sum(times) / len(times)
    kmers = sample * (L - k + 1)
    val = kmers / dt
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": min(a.warmup, 1),
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "impl": "reference", "config": workload(a),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": "%d reads (%d k-mers) of the workload per step; runs through FileDump files in %s, "
                                       "8-thread chunk workers like KMerCounter.cpp:117, single KMerFileMerger" %
                                       (sample, kmers, tmp_root or "tmp")},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline(a):
    import oracle
    cores = os.cpu_count() or 1
    L, k = a.read_len, a.k
    sample = min(a.reads, a.cpu_sample_reads)
    kind = "reference" if oracle.ref_available() else "port"
    reads = oracle.gen_reads(sample, L, a.genome, a.sub_rate, a.n_rate, seed=a.seed, zipf_loci=a.zipf_loci)
    tmp_root = "/dev/shm" if os.path.isdir("/dev/shm") else None
    t0 = time.perf_counter()
    if kind == "reference":
        with tempfile.TemporaryDirectory(dir=tmp_root) as d:
            oracle.ref_count_packed(reads, L, k, 89364, cores, d, os.path.join(d, "out.bin"))
    else:
        oracle.count(reads, L, k, 89364, threads=cores)
    dt = time.perf_counter() - t0
    kmers = sample * (L - k + 1)
    return {"value": kmers / dt, "unit": UNIT, "cores": cores, "kind": kind, "seconds": dt,
            "sample": "first %d reads (%d k-mers) of the workload, chunks of 89,364 reads (the reference's default), "
                      "%d chunk-worker threads, one serial KMerFileMerger" % (sample, kmers, cores)}


# ------------------------------------------------------------------- our arm
def run_ours(a, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    import kmer_counter_b200 as kc
    from kmer_counter_b200 import multigpu, synth

    numa = bind_near_gpu(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L, k, R = a.read_len, a.k, a.reads
    nk = L - k + 1
    n_bytes = R * L
    # high priority: the combine step and the record read-back of chunk i are queued while the
    # kernels of chunks i+1, i+2 occupy the SMs on the (default-priority) slot streams
    stream = torch.cuda.Stream(device=dev, priority=-1)
    torch.cuda.set_stream(stream)
    counter = kc.Counter(k, L, device=local_rank, method=a.method, n_slots=E2E_SLOTS, max_chunk_bytes=0 if a.no_e2e else n_bytes, table_slots=a.target,
                         stream=stream.cuda_stream)
    d_reads = torch.empty(n_bytes + 256, dtype=torch.uint8, device=dev)
    synth.synth_reads_device(d_reads.data_ptr(), R, L, a.genome, a.sub_rate, a.n_rate, a.seed,
                             first_read=rank * R, zipf_loci=a.zipf_loci, stream=stream.cuda_stream)
    stream.synchronize()

    if a.check and rank == 0:
        import oracle
        pre = min(R, 200_000)
        host = d_reads[: pre * L].cpu().numpy()
        assert bytes(host) == bytes(oracle.gen_reads(pre, L, a.genome, a.sub_rate, a.n_rate, seed=a.seed,
                                                     zipf_loci=a.zipf_loci)), "device generator != host generator"
        want = oracle.count(host, L, k, threads=os.cpu_count() or 1)
        r = counter.count_device(d_reads.data_ptr(), pre * L)
        assert r.to_bytes() == want, "prefix parity failed"
        r.free()
        sys.stderr.write("check: %d-read prefix bit-exact against the oracle\n" % pre)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peer = None
    if world > 1 and os.environ.get("KC_NO_PEER") != "1":
        try:            # exchange fused into the combine kernel over NVLink peer memory
            peer = multigpu.PeerCombine(counter, dev, max_records=int(R * nk * 0.45))
        except Exception as e:
            sys.stderr.write("peer-memory exchange unavailable (%s): using NCCL\n" % e)
            peer = None

    def step():
        if world > 1:
            run = multigpu.count_shard(counter, d_reads.data_ptr(), n_bytes, dev, peer=peer)
        else:
            run = counter.count_device(d_reads.data_ptr(), n_bytes)
        n = len(run)
        run.free()
        return n

    for _ in range(a.warmup):
        step()
    barrier()
    st0 = counter.stats()
    dom_ms, dom_bytes, dom_launch, tot_ms = 0.0, 0, 0, 0.0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        ev0.record(stream)
        t0 = time.perf_counter()
        distinct = 0
        for _ in range(a.steps):
            distinct = step()
            s = counter.stats()
            dom_ms += s["ms_dominant"]; dom_bytes += s["dominant_bytes"]; dom_launch += s["dominant_launches"]
            tot_ms += s["ms_total"]
        ev1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    st1 = counter.stats()
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / a.steps
    kmers_step = R * nk * world
    value = kmers_step / (ms_per_step * 1e-3)
    method_used = st1["method_used"]
    launches = st1["launches"] - st0["launches"]

    # ---- end to end through the host-buffer API (pinned in, pinned out)
    e2e = None
    if not a.no_e2e:
        host_reads = d_reads[:n_bytes].cpu().numpy()
        for sl in range(E2E_SLOTS):                         # every pinned input slot holds the step's reads
            counter.slot_buffer(sl)[:n_bytes] = host_reads
        out_cap = (distinct + 1024) * counter.record_size if world == 1 else (R * nk // 2) * counter.record_size
        pinned_out = counter.host_alloc(out_cap)
        d2h = 0
        e2e_parts = {"h2d_count": 0.0, "exchange_merge": 0.0, "records_d2h": 0.0}

        def finish(sl):
            """wait for the slot's chunk, [exchange + combine], records D2H into pinned memory"""
            t_a = time.perf_counter()
            run = counter.wait(sl)
            t_b = time.perf_counter()
            if world > 1:
                merged = peer.combine(run) if peer is not None else None
                run = merged if merged is not None else multigpu.exchange_and_combine(counter, run, dev)
            t_c = time.perf_counter()
            nb = run.copy_into(pinned_out.ctypes.data, out_cap)
            run.free()
            t_d = time.perf_counter()
            e2e_parts["h2d_count"] += t_b - t_a; e2e_parts["exchange_merge"] += t_c - t_b; e2e_parts["records_d2h"] += t_d - t_c
            return nb

        # (1) one step at a time: the latency of a single chunk, with its parts
        for _ in range(2):
            counter.submit(0, n_bytes); finish(0)
        for kx in e2e_parts:
            e2e_parts[kx] = 0.0
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            counter.submit(0, n_bytes); d2h = finish(0)
        barrier()
        single_ms = (time.perf_counter() - t0) / a.steps * 1e3
        parts = {kx: v / a.steps * 1e3 for kx, v in e2e_parts.items()}
        # (2) the reported number: the same K steps pipelined over the pinned slots, which is how the
        # API is meant to be driven (kc_submit / kc_wait from the producer, the records read back by a
        # consumer thread): H2D of step i, the kernels of step i-1 and the records D2H of step i-2
        # overlap (two copy engines + SMs). Every step still moves all its bytes both ways.
        import queue
        todo, res = queue.Queue(maxsize=1), {"nb": 0, "err": None}

        trace = [] if os.environ.get("KC_E2E_TRACE") == "1" and rank == 0 else None   # (what, start ms, ms) per call

        def timed(what, fn, *args):
            if trace is None:
                return fn(*args)
            t_s = time.perf_counter()
            out = fn(*args)
            trace.append((what, t_s * 1e3, (time.perf_counter() - t_s) * 1e3))
            return out

        def reader():
            while True:
                run = todo.get()
                if run is None:
                    return
                try:
                    res["nb"] = timed("copy", run.copy_into, pinned_out.ctypes.data, out_cap)
                    run.free()
                except Exception as e:                       # surfaced after the join
                    res["err"] = e

        def combine(run):
            merged = peer.combine(run) if peer is not None else None
            return merged if merged is not None else multigpu.exchange_and_combine(counter, run, dev)

        def emit(run):
            """[exchange + combine of a finished chunk], then hand its records to the reader"""
            if world > 1:
                run = timed("combine", combine, run)
            timed("put", todo.put, run)

        def pipelined(n_steps):
            # Steady state of iteration i: H2D of chunk i is in flight, chunk i-1 is being counted, chunk
            # i-2 is combined with the peers' parts as soon as i-1's kernels have drained (so the combine
            # runs in the shadow of i's H2D instead of fighting i-1's kernels for SMs), and the reader
            # thread copies the records of chunk i-3 to the host.
            th = threading.Thread(target=reader, daemon=True)
            th.start()
            held = None                                     # counted, not yet combined
            for i in range(n_steps + 1):
                if i < n_steps:
                    timed("submit%d" % (i % E2E_SLOTS), counter.submit, i % E2E_SLOTS, n_bytes)
                if i >= 1:
                    sl = (i - 1) % E2E_SLOTS
                    run = timed("wait%d" % sl, counter.wait, sl)
                    if world == 1:                          # nothing to combine: straight to the reader
                        emit(run)
                        continue
                    if held is not None:
                        emit(held)
                    held = run
            if held is not None:
                emit(held)
            todo.put(None)
            th.join()
            if res["err"] is not None:
                raise res["err"]

        pipelined(max(a.warmup, 2 * E2E_SLOTS))             # untimed: every slot's arena and the run pool reach their steady size
        barrier()
        t0 = time.perf_counter()
        if trace is not None:
            del trace[:]
        pipelined(a.steps)
        d2h = res["nb"]
        barrier()
        if trace:
            t_first = min(t for _, t, _ in trace)
            for what, t_s, dur in sorted(trace, key=lambda x: x[1]):
                sys.stderr.write("e2e-trace %8.1f  %-9s %7.2f\n" % (t_s - t_first, what, dur))
        e2e_dt = (time.perf_counter() - t0) / a.steps
        t = torch.tensor([e2e_dt, single_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt, single_ms = float(t[0].item()), float(t[1].item())
        e2e = {"value": kmers_step / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": n_bytes, "d2h_bytes_per_step": int(d2h),
               "ms_per_step": e2e_dt * 1e3,
               "timing": "wall clock around K steps pipelined over %d pinned slots (kc_submit/kc_wait, records read back "
                         "by a consumer thread), sync on both sides, max over ranks; every step copies its reads H2D "
                         "and its records D2H" % E2E_SLOTS,
               "single_step_ms": single_ms, "single_step_value": kmers_step / (single_ms * 1e-3),
               "single_step_parts_ms_rank0": parts, "cpus_bound_near_gpu": numa}
        counter.host_free(pinned_out)

    clocks = clk.summary()
    peak, peak_src = measured_peak_gbs()
    if rank == 0:
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        dom_name = st1["stage_names"][st1["dominant_stage"]] if st1["stage_names"] else "none"
        # DRAM traffic of the dominant kernel from the committed ncu --set full capture of this
        # workload (bytes per launch), if there is one for this method and size
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            ent = tj.get("%s/%s/%d" % (method_used, dom_name, R))
            if ent:
                traffic, traffic_src = ent["dram_bytes_per_launch"], ent["source"]
        except Exception:
            pass
        Kb, S = 8 * counter.words, counter.record_size
        b_alg = R * L + R * nk * (Kb + 8) + distinct * S
        path_ach = b_alg / (tot_ms / a.steps * 1e-3) / 1e9 if tot_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": dict(workload(a), method=method_used,
                                                                     exchange=None if world == 1 else
                                                                     "peer memory: combine kernel loads the parts from the peers' HBM over NVLink (CUDA IPC)"
                                                                     if peer is not None else "NCCL grouped send/recv, then combine"),
            "bases_per_s": R * L * world / (ms_per_step * 1e-3),
            "distinct_per_gpu": distinct,
            "roofline": {"bound": "hbm", "kernel": dom_name,
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                         "peak_source": peak_src + " HBM copy bandwidth (MEASURED_PEAKS.json)",
                         "bytes_per_launch": dom_bytes / max(dom_launch, 1), "launches_per_step": dom_launch / a.steps,
                         "ms_per_launch": dom_ms / max(dom_launch, 1),
                         "share_of_step": dom_ms / tot_ms if tot_ms else None, "traffic": traffic,
                         "traffic_source": traffic_src},
            "path_roofline": {"bound": "hbm", "b_alg_bytes": b_alg, "achieved": path_ach, "peak": peak, "unit": "GB/s",
                              "frac": path_ach / peak if peak else None,
                              "definition": "SURVEY 8(d): B_in + N*(Kb+8) + U*S over the local counting time"},
            "stages_last_step": {n: round(m, 4) for n, m in zip(st1["stage_names"], st1["ms_stage"])},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "wall_ms_per_step": wall / a.steps * 1e3,
        }
        if not a.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(a)
        print(json.dumps(line), flush=True)
    counter.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank)
        return
    if world != a.gpus and world == 1 and a.gpus > 1:
        sys.stderr.write("bench.py: --gpus %d needs torchrun (WORLD_SIZE=%d)\n" % (a.gpus, world))
        sys.exit(2)
    run_ours(a, rank, world, local_rank)


if __name__ == "__main__":
    main()

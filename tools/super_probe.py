"""GPU probe for the super-window path: parity cases against the oracle with the pipeline's
device scalars printed, then a timed C2-shaped run. Development tool (gpurun), not a test."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import kmer_counter_b200 as kc  # noqa: E402


def diff(got, want, S):
    g = np.frombuffer(got, dtype=np.uint8).reshape(-1, S) if len(got) % S == 0 else None
    w = np.frombuffer(want, dtype=np.uint8).reshape(-1, S)
    if g is None:
        return "got %d bytes (not a multiple of %d)" % (len(got), S)
    msg = "records got %d want %d" % (len(g), len(w))
    n = min(len(g), len(w))
    neq = np.nonzero((g[:n] != w[:n]).any(axis=1))[0]
    if len(neq):
        i = int(neq[0])
        msg += "; first diff at %d: got %s want %s" % (i, g[i].tobytes().hex(), w[i].tobytes().hex())
        gs = sum(int.from_bytes(r[-4:].tobytes(), "little") for r in g)
        ws = sum(int.from_bytes(r[-4:].tobytes(), "little") for r in w)
        msg += "; count sums got %d want %d" % (gs, ws)
    return msg


def case(R, L, k, G, e, n, seed=None, compat="ref", reads=None, **kw):
    if reads is None:
        reads = oracle.gen_reads(R, L, G, e, n, seed=seed if seed is not None else R + k)
    if compat == "ref":
        want = oracle.process_chunk(reads, L, k)
    else:
        want = oracle.naive_count(reads, L, k, strict=True)
    with kc.Counter(k, L, method="super", compat=compat, n_slots=2, max_chunk_bytes=max(len(reads), 1 << 20), **kw) as c:
        got = c.process_chunk(reads)
        st = c.stats()
        sc = c.debug_scalars()
    ok = got == want
    print("%s R=%d L=%d k=%d G=%d e=%g n=%g %s method=%s %s" % (
        "ok  " if ok else "FAIL", len(reads) // L, L, k, G, e, n, kw or "", st["method_used"],
        "" if ok else diff(got, want, 8 * ((k + 31) // 32) + 4)), flush=True)
    if not ok or os.environ.get("VERBOSE"):
        print("     scalars", sc, flush=True)
    return ok


def accum_cases():
    """accumulating mode: chunks through the pinned slots / device buffers, one count"""
    import torch
    ok = True
    for (R, L, k, G, e, n, expected, chunk_reads) in [
            (30000, 100, 31, 200000, 0.01, 0.002, 30000, 7000),     # fits the plan: one count
            (30000, 100, 31, 200000, 0.01, 0.002, 7000, 7000),      # planned too small: parts + merge
            (20000, 100, 63, 100000, 0.001, 0.001, 20000, 4096),
            (5000, 70, 28, 0, 0.0, 0.01, 0, 1000)]:
        reads = oracle.gen_reads(R, L, G, e, n, seed=R + k + 1)
        want = oracle.count(reads, L, k)
        for how in ("slots", "device"):
            with kc.Counter(k, L, method="super", n_slots=2, max_chunk_bytes=chunk_reads * L) as c:
                c.accum_begin(expected)
                if how == "slots":
                    sl = 0
                    busy = [False, False]
                    for r0 in range(0, R, chunk_reads):
                        part = reads[r0 * L:(r0 + chunk_reads) * L]
                        if busy[sl]:
                            c.accum_wait(sl)
                        c.slot_buffer(sl)[:len(part)] = part
                        c.accum_submit(sl, len(part))
                        busy[sl] = True
                        sl ^= 1
                else:
                    d = torch.from_numpy(reads.copy()).cuda()
                    c.accum_add_device(d.data_ptr(), len(reads))
                run = c.accum_flush()
                got = run.to_bytes()
                run.free()
                sc = c.debug_scalars()
            good = got == want
            ok &= good
            print("%s accum %s R=%d L=%d k=%d expected=%d chunk=%d %s" % (
                "ok  " if good else "FAIL", how, R, L, k, expected, chunk_reads,
                "" if good else diff(got, want, 8 * ((k + 31) // 32) + 4)), flush=True)
            if not good:
                print("     scalars", sc, flush=True)
    return ok


def xchg_cases():
    """multi-GPU exchange with all ranks as contexts of this process on device 0 (kc_xchg_run_all)"""
    import torch
    from kmer_counter_b200 import engine
    ok = True
    for (P, R, L, k, G, e, n) in [(2, 20000, 100, 31, 100000, 0.01, 0.002), (4, 30000, 100, 31, 50000, 0.01, 0.001),
                                  (8, 40000, 100, 31, 0, 0.0, 0.0), (3, 9000, 100, 63, 40000, 0.001, 0.001),
                                  (8, 800, 100, 31, 3000, 0.0, 0.01), (4, 12000, 70, 28, 30000, 0.01, 0.001)]:
        reads = oracle.gen_reads(R, L, G, e, n, seed=P * 1000 + k)
        want = oracle.count(reads, L, k)
        per = (R + P - 1) // P
        cs = [kc.Counter(k, L, method="super") for _ in range(P)]
        try:
            bufs = []
            for r, c in enumerate(cs):
                c.xchg_begin(r, P, per)
                part = reads[r * per * L:(r + 1) * per * L]
                d = torch.from_numpy(part.copy()).cuda() if len(part) else torch.empty(16, dtype=torch.uint8, device="cuda")
                bufs.append(d)
                c.accum_add_device(d.data_ptr(), len(part))
            runs = engine.xchg_run_all(cs)
            got = b"".join(r.to_bytes() for r in runs)
            sizes = [len(r) for r in runs]
            for r in runs:
                r.free()
        finally:
            for c in cs:
                c.close()
        good = got == want
        ok &= good
        print("%s xchg P=%d R=%d L=%d k=%d sizes=%s %s" % ("ok  " if good else "FAIL", P, R, L, k, sizes,
              "" if good else diff(got, want, 8 * ((k + 31) // 32) + 4)), flush=True)
    return ok


def main():
    ok = True
    ok &= case(2000, 100, 31, 30000, 0.0, 0.0)
    ok &= case(2000, 100, 31, 30000, 0.01, 0.002)
    ok &= case(1500, 100, 32, 20000, 0.01, 0.001)
    ok &= case(1500, 100, 28, 20000, 0.0, 0.001)
    ok &= case(1500, 100, 29, 20000, 0.0, 0.001)
    ok &= case(700, 70, 63, 5000, 0.001, 0.001)
    ok &= case(700, 70, 60, 5000, 0.001, 0.001)
    ok &= case(500, 41, 33, 0, 0.0, 0.01)
    ok &= case(257, 33, 33, 0, 0.0, 0.0)
    ok &= case(1000, 133, 31, 8000, 0.002, 0.0)
    ok &= case(6000, 100, 31, 0, 0.0, 0.001, seed=77)                       # iid: all distinct
    ok &= case(6000, 100, 31, 0, 0.0, 0.001, seed=77, table_slots=64)       # tiny bins
    ok &= case(6000, 100, 31, 0, 0.0, 0.001, seed=77, table_slots=100000)   # bins beyond the table: split passes
    ok &= case(2000, 100, 31, 30000, 0.01, 0.002, compat="strict")
    L = 100
    hot = oracle.gen_reads(3, L, 0, 0, 0, seed=5)
    for reps, kk in ((4000, 31), (30000, 31), (30000, 63)):
        reads = np.concatenate([np.tile(hot, reps), oracle.gen_reads(2000, L, 50000, 0.01, 0.001, seed=6)])
        ok &= case(0, L, kk, 0, 0, 0, reads=reads)
    reads = np.frombuffer((b"A" * 69) * 30 + (b"T" * 69) * 20 + (b"A" * 40 + b"N" + b"A" * 28) * 3, dtype=np.uint8)
    ok &= case(0, 69, 31, 0, 0, 0, reads=reads)
    ok &= case(0, 50, 31, 0, 0, 0, reads=np.frombuffer(b"N" * (50 * 40), dtype=np.uint8))
    ok &= case(100000, 100, 31, 1000000, 0.0, 0.001, seed=1)                # C1
    ok &= case(200000, 100, 63, 2000000, 0.001, 0.0, seed=3)
    os.environ["KC_SW_FORCE_DUP"] = "1"
    print("forced folding path:")
    ok &= case(2000, 100, 31, 30000, 0.01, 0.002)
    ok &= case(100000, 100, 31, 1000000, 0.0, 0.001, seed=1)
    ok &= case(700, 70, 63, 5000, 0.001, 0.001)
    os.environ["KC_SW_FORCE_DUP"] = "0"
    ok &= accum_cases()
    ok &= xchg_cases()
    print("ALL OK" if ok else "SOME FAILED", flush=True)

    # timing, C2 shape
    if len(sys.argv) > 1 and sys.argv[1] == "time":
        import torch
        from kmer_counter_b200 import synth
        R, L, k = 10_000_000, 100, 31
        dev = torch.device("cuda", 0)
        d_reads = torch.empty(R * L + 256, dtype=torch.uint8, device=dev)
        synth.synth_reads_device(d_reads.data_ptr(), R, L, 100_000_000, 1e-3, 0.0, 2)
        torch.cuda.synchronize()
        for method in ("super", "hash"):
            with kc.Counter(k, L, method=method) as c:
                for i in range(4):
                    t0 = time.perf_counter()
                    run = c.count_device(d_reads.data_ptr(), R * L)
                    dt = time.perf_counter() - t0
                    n = len(run)
                    run.free()
                st = c.stats()
                print(method, "records", n, "wall ms %.2f" % (dt * 1e3), "ms_total %.2f" % st["ms_total"],
                      dict(zip(st["stage_names"], [round(x, 3) for x in st["ms_stage"]])), flush=True)
                if method == "super":
                    print("   scalars", c.debug_scalars(), flush=True)
        # digest check between the two methods on the full C2 input
        import hashlib
        digs = []
        for method in ("super", "hash"):
            with kc.Counter(k, L, method=method) as c:
                run = c.count_device(d_reads.data_ptr(), R * L)
                digs.append(hashlib.sha256(run.to_bytes()).hexdigest())
                run.free()
        print("C2 full-size super == hash:", digs[0] == digs[1], digs[0][:16], flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())

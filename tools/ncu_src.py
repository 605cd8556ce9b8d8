"""Per-source-line view of an `ncu --page source --csv --print-source cuda,sass` dump for ONE kernel.
usage: ncu_src.py src.csv <kernel-substring> [top] [inst|samp]"""
import csv, sys
rows = csv.reader(open(sys.argv[1]))
want = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
order = sys.argv[4] if len(sys.argv) > 4 else "samp"
keep, cur, hdr, out = False, None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        keep = want in r[1]
    elif r[0] == "Line No":
        hdr = r
    elif keep and hdr and r[0].isdigit():
        n = len(r) - len(hdr)            # commas inside the source text shift the columns
        g = lambda name: r[hdr.index(name) + n]
        try:
            out.append((cur, int(r[0]), ",".join(r[1:2 + n])[:70], int(g("# Samples")), int(g("Instructions Executed")),
                        int(g("Thread Instructions Executed")), int(g("stall_barrier")), int(g("stall_short_sb")),
                        int(g("stall_long_sb")), int(g("stall_wait")), int(g("L1 Wavefronts Shared Excessive") or 0)))
        except ValueError:
            pass
ts, ti, tt = sum(o[3] for o in out) or 1, sum(o[4] for o in out) or 1, sum(o[5] for o in out) or 1
print("kernel ~%s: %d samples, %d warp-inst, %.1f lanes/inst" % (want, ts, ti, tt / ti))
key = (lambda o: -o[4]) if order == "inst" else (lambda o: -o[3])
print("%6s %6s %5s %5s %5s %5s %5s  %s" % ("samp%", "inst%", "lanes", "bar", "ssb", "lsb", "wait", "line"))
for o in sorted(out, key=key)[:top]:
    print("%5.1f%% %5.1f%% %5.1f %5d %5d %5d %5d  %s:%d %s" % (100 * o[3] / ts, 100 * o[4] / ti, o[5] / max(o[4], 1), o[6], o[7], o[8], o[9], o[0], o[1], o[2]))

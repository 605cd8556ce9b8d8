"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by source line."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
want = sys.argv[3] if len(sys.argv) > 3 else None      # substring of the kernel name to keep
keep = True
cur_file, hdr = None, None
agg = collections.OrderedDict()
line_key = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name':
        keep = want is None or want in r[1]
        continue
    if not keep: continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None: continue
    if r[0] != '':          # a source line row: the text may contain commas -> columns shift; take line no + text
        line_key = (cur_file, int(r[0]) if r[0].isdigit() else -1, r[1][:80])
        agg.setdefault(line_key, [0, 0])
        continue
    # sass row: columns aligned with hdr
    try:
        i_inst = hdr.index('Instructions Executed'); i_samp = hdr.index('# Samples')
        inst = int(r[i_inst]) if r[i_inst].isdigit() else 0
        samp = int(r[i_samp]) if r[i_samp].isdigit() else 0
    except Exception:
        continue
    if line_key: agg[line_key][0] += inst; agg[line_key][1] += samp
tot_i = sum(v[0] for v in agg.values()) or 1; tot_s = sum(v[1] for v in agg.values()) or 1
print("total warp-inst %d, samples %d" % (tot_i, tot_s))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% samp %5.1f%% inst  %s:%d  %s" % (100 * v[1] / tot_s, 100 * v[0] / tot_i, k[0], k[1], k[2]))

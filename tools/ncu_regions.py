"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by 10-line source regions of one kernel:
share of warp instructions, average active lanes, share of stall samples."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]
step = int(sys.argv[3]) if len(sys.argv) > 3 else 10
keep=False; hdr=None; agg=collections.OrderedDict(); line_key=None; cur_file=None
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if r[0]=='Function Name': keep=want in r[1]; continue
    if not keep: continue
    if r[0]=='Line No': hdr=r; continue
    if hdr is None: continue
    if r[0]!='':
        line_key=(cur_file,int(r[0]) if r[0].isdigit() else -1); agg.setdefault(line_key,[0,0,0]); continue
    try:
        ii=hdr.index('Instructions Executed'); inst=int(r[ii]) if r[ii].isdigit() else 0
        it=hdr.index('Thread Instructions Executed'); ti=int(r[it]) if r[it].isdigit() else 0
        isamp=hdr.index('# Samples'); sa=int(r[isamp]) if r[isamp].isdigit() else 0
    except Exception as e: continue
    if line_key: agg[line_key][0]+=inst; agg[line_key][1]+=ti; agg[line_key][2]+=sa
tot=sum(v[0] for v in agg.values()) or 1; tots=sum(v[2] for v in agg.values()) or 1
print('total warp-inst',tot)
b=collections.Counter(); bt=collections.Counter(); bs=collections.Counter()
for (f,l),v in agg.items():
    key = (f, (l//step)*step) if f.startswith('kc_') else (f,0)
    b[key]+=v[0]; bt[key]+=v[1]; bs[key]+=v[2]
for k in sorted(b):
    if b[k]/tot>0.008 or bs[k]/tots>0.01: print("%-36s %5.1f%% inst  lanes %4.1f  %5.1f%% samp"%(k, 100*b[k]/tot, bt[k]/max(b[k],1), 100*bs[k]/tots))

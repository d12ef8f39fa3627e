"""Aggregate `ncu --page source --print-source cuda,sass --csv` output by kernel and source line:
   python tools/ncu_lines.py report.csv > lines.txt   (samples and executed warp instructions)"""
import csv, sys, collections
rows = csv.reader(open(sys.argv[1], newline=""))
kern = None; cur = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, ""])
stall = collections.defaultdict(lambda: collections.Counter())
for r in rows:
    if len(r) >= 2 and r[0] == "Function Name": kern = r[1].split("(")[0].split("::")[-1][:24]; continue
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] != "":
        try: ln = int(r[0])
        except ValueError: continue
        def num(x):
            try: return int(float(x))
            except ValueError: return 0
        a = agg[(kern, cur, ln)]
        a[0] += num(r[6]); a[1] += num(r[7]); a[2] = r[1].strip()[:110]
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h:
                v = num(r[i])
                if v: stall[(kern, cur, ln)][h[6:]] += v
for k in sorted(set(x[0] for x in agg)):
    tot = sum(v[0] for kk, v in agg.items() if kk[0] == k)
    ins = sum(v[1] for kk, v in agg.items() if kk[0] == k)
    print("=== %s: %d samples, %d warp instructions" % (k, tot, ins))
    for kk, v in sorted(agg.items(), key=lambda kv: (kv[0][1], kv[0][2])):
        if kk[0] != k or (v[0] == 0 and v[1] < 30): continue
        top = ",".join("%s:%d" % (a, b) for a, b in stall[kk].most_common(3))
        print("%6d smp %7d ins  %-14s:%4d  %-40s | %s" % (v[0], v[1], kk[1], kk[2], top, v[2]))

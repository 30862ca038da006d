"""Top stall lines of an `ncu --page source --csv` dump: python tools/ncu_src_top.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
idx = {n: i for i, n in enumerate(hdr)}
stall_cols = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[idx["# Samples"]] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
agg = {}
for n, r in enumerate(body):
    s = int(r[idx["# Samples"]] or 0)
    body[n] = (s, n, r)
for s, n, r in sorted(body, key=lambda t: -t[0])[:top]:
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:3]
    print("%5d %5.1f%% #%4d ex=%8s %-60s %s" % (s, 100.0 * s / tot, n, r[idx["Instructions Executed"]], r[idx["Source"]].strip()[:60],
                                        " ".join("%s:%d" % (b, a) for a, b in st if a)))

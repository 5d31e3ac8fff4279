"""Top stall locations of an `ncu --page source --csv` export (SASS view): python tools/ncu_top.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
tot_inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot, "instructions executed", tot_inst)
agg = {s: sum(int(r[ix[s]] or 0) for r in body) for s in stalls}
print("stall mix:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:n]
for i in sorted(order):
    r = body[i]
    top = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print("%5d %5.2f%% exec %9s  %-70s %s" % (i, 100.0 * int(r[ix["# Samples"]] or 0) / max(tot, 1), r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:70],
                                           " ".join("%s:%d" % (b, a) for a, b in top if a)))

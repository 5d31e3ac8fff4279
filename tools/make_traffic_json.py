"""profiles/traffic.json from tools/measure_traffic.sh's outputs: python tools/make_traffic_json.py gpurun_out c2 c3 c4"""
import csv, json, os, sys
d, out = sys.argv[1], {}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
try:
    out = json.load(open(path))
    if "dram_bytes_per_position" in out:
        out = {}
except Exception:
    out = {}
for w in sys.argv[2:]:
    rows = [r for r in csv.reader(open(os.path.join(d, "traffic_%s.csv" % w))) if len(r) > 10]
    hdr = rows[0]
    m = {}
    for r in rows[1:]:
        rec = dict(zip(hdr, r))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
        m[rec["Metric Name"]] = float(rec["Metric Value"].replace(",", "")) * scale.get(rec["Metric Unit"], 1.0)
        kernel = rec["Kernel Name"]
    steps = json.load(open(os.path.join(d, "steps_%s.json" % w)))
    pos, ms, launches = steps["per_step"][0]  # the first timed step = the 4th launch (3 warm-up steps)
    assert launches == 1
    dram = m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
    out[w] = {"dram_bytes_per_position": dram / pos, "dram_bytes_read": m["dram__bytes_read.sum"],
              "dram_bytes_write": m["dram__bytes_write.sum"], "positions_in_launch": pos,
              "launch_ms_under_ncu": m.get("gpu__time_duration.sum", 0) * 1e3, "launch_ms_in_bench": ms, "kernel": kernel,
              "source": "tools/measure_traffic.sh: bench.py --workload %s under ncu (one pass), first timed step" % w}
    print(w, out[w])
json.dump(out, open(path, "w"), indent=1)

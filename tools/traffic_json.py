"""gpurun_out/traffic_<tag>.csv (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over one full-size training step)
-> profiles/ncu_traffic.json (feeds roofline.traffic in bench.py).   python tools/traffic_json.py <csv> <source note>"""
import collections, csv, io, json, sys
src, note = sys.argv[1], sys.argv[2]
lines = [l for l in open(src) if not l.startswith("==")]
agg = collections.OrderedDict()
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
ids = {}
for row in csv.DictReader(io.StringIO("".join(lines))):
    if not row.get("Metric Name", "").startswith("dram__bytes"):
        continue
    name = row["Kernel Name"].split("(")[0].replace("papr::", "").strip()
    v = float(row["Metric Value"].replace(",", "")) * scale[row["Metric Unit"]]
    a = agg.setdefault(name, {"launches": 0, "dram_bytes": 0.0})
    a["dram_bytes"] += v
    ids.setdefault(name, set()).add(row["ID"])
for k in agg:
    agg[k]["launches"] = len(ids[k])
out = {"source": note, "per_kernel": agg,
       "tcgen05_kernels_bytes_per_step": sum(v["dram_bytes"] for v in agg.values())}
json.dump(out, open("profiles/ncu_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))

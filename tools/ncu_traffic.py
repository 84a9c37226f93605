"""Extract per-kernel DRAM traffic / duration of one full-load round from an `ncu --page raw --csv` dump into the small JSON
bench.py reads for its `traffic` fields (profiles/ncu_traffic_r02.json).  usage: ncu_traffic.py raw.csv out.json"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
out = {"source": sys.argv[1].split("/")[-1], "launches": []}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("obca::", "")
    rd = float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]]
    wr = float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]]
    t = float(d["gpu__time_duration.sum"]) * tscale[u["gpu__time_duration.sum"]]
    out["launches"].append({"kernel": name, "dram_read_bytes": rd, "dram_write_bytes": wr, "time_us": t})
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out["launches"]))

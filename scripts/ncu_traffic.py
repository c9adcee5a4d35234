"""ncu raw page (csv) of scripts/ncu_dominant.py -> profiles/r2_ncu_traffic.json: DRAM bytes per launch of the dominant kernel
(dram__bytes_read.sum + dram__bytes_write.sum, averaged over the captured launches).
Usage: python scripts/ncu_traffic.py gpurun_out/r2_ncu_raw_dominant.csv"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    if "conv_tc2_kernel" not in name:
        continue
    rd = float(r[ix["dram__bytes_read.sum"]]) * scale[units[ix["dram__bytes_read.sum"]]]
    wr = float(r[ix["dram__bytes_write.sum"]]) * scale[units[ix["dram__bytes_write.sum"]]]
    key = name.split("(")[0].replace("void ", "").replace("rave::tc::", "").replace("tc::", "")
    out.setdefault(key, []).append((rd, wr, float(r[ix["gpu__time_duration.sum"]])))
res = {}
for k, v in out.items():
    res[k] = {"launches": len(v), "dram_read_bytes": sum(a for a, _, _ in v) / len(v),
              "dram_write_bytes": sum(b for _, b, _ in v) / len(v), "traffic_bytes": sum(a + b for a, b, _ in v) / len(v),
              "ncu_duration_us": sum(c for _, _, c in v) / len(v),
              "layer": "MSD scale-0 layer 3: B=64, 192 -> 384 channels, Lin 4096 -> Lout 1024, k15 s4 (scripts/ncu_dominant.py)",
              "source": sys.argv[1]}
json.dump(res, open("profiles/r2_ncu_traffic.json", "w"), indent=1)
print(json.dumps(res, indent=1))

"""profiles/r02_ncu_traffic.json from the raw export of an `ncu --set full` capture of
scripts/profile_kernels.py: DRAM bytes per launch of every hot-path kernel (bench.py's
roofline.traffic reads this file).   python scripts/ncu_traffic.py <prof_raw.csv> <source note>"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
kn, rd, wr, tm = (hdr.index(x) for x in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                         "gpu__time_duration.sum"))
U = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
names = {"ch_rhs_tma_kernel<4, 4, 0, 168, 0>": "ch_rhs_kernel", "fft_chain_kernel<0, 0, 3, 2>": "fft_zy_forward",
         "fft_chain_kernel<1, 0, 3, 3>": "fft_yz_inverse+u", "fft_line_ws_kernel<StridedLine<512, 8, 2>, 3>": "fft_x_fwd*filter*inv"}
out = {"size": 512, "source": sys.argv[2] if len(sys.argv) > 2 else sys.argv[1], "kernels": {}}
val = lambda r, i: float(r[i].replace(",", "")) * U[units[i]]
for r in rows[2:]:
    for key, bn in names.items():
        if key in r[kn] and key not in out["kernels"]:
            out["kernels"][key] = {"bench_name": bn, "dram_bytes": val(r, rd) + val(r, wr), "dram_read_bytes": val(r, rd),
                                   "dram_write_bytes": val(r, wr), "ncu_duration": r[tm] + " " + units[tm]}
out["step_dram_bytes"] = sum(v["dram_bytes"] for v in out["kernels"].values())
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out["kernels"], indent=1))

"""Summarise an `ncu --page raw --csv` export (run here, no GPU needed):
    ncu -i gpurun_out/X_prof.ncu-rep --page raw --csv > gpurun_out/X_raw.csv
    python tools/ncu_summary.py gpurun_out/X_raw.csv profiles/X_ncu_summary.txt [profiles/ncu_traffic.json [arithmetic]]
Writes a per-launch table (duration, DRAM bytes, DRAM / tensor / SM utilisation, registers, grid) and, optionally, the per-kernel
DRAM traffic (bytes per launch, averaged over the captured launches) that bench.py reports as roofline.traffic."""
import csv
import json
import re
import sys

KIND = [("edge_aggregate", "edge_aggregate"), ("tc_edge_encoder", "edge_encoder"), ("tc_node_encoder", "node_encoder"),
        ("tc_node_update_kernel<(bool)0>", "node_update"), ("tc_node_update_kernel<(bool)1>", "node_update_head"),
        ("tc_node_update_kernel<0>", "node_update"), ("tc_node_update_kernel<1>", "node_update_head"),
        ("tc_node_update_kernel<0,", "node_update"), ("tc_node_update_kernel<1,", "node_update_head"), ("knn_rows", "graph_knn_rows")]
COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdGB"), ("dram__bytes_write.sum", "wrGB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]


def to_gb(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}[unit]


def main():
    raw, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {k: i for i, k in enumerate(hdr)}
    lines = [f"# {raw}: ncu --set full --clock-control none, one line per captured launch (cold-cache, serialised replays)",
             f"{'kernel':40s}" + "".join(f"{n:>10s}" for _, n in COLS)]
    traffic = {}
    for r in data:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("agx::", "").replace("void ", "")
        vals = []
        for k, n in COLS:
            v, u = r[ix[k]], units[ix[k]]
            if n in ("rdGB", "wrGB"):
                vals.append(f"{to_gb(v, u):10.3f}")
            elif n == "us":
                f = float(v.replace(",", "")) * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
                vals.append(f"{f:10.1f}")
            else:
                vals.append(f"{float(v.replace(',', '')):10.1f}")
        lines.append(f"{name[:40]:40s}" + "".join(vals))
        full = r[ix["Kernel Name"]]
        for pat, kind in KIND:
            if pat in full:
                t = traffic.setdefault(kind, {"launches": 0, "dram_read_gb": 0.0, "dram_write_gb": 0.0})
                t["launches"] += 1
                t["dram_read_gb"] += to_gb(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
                t["dram_write_gb"] += to_gb(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
                break
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if len(sys.argv) > 3:
        for t in traffic.values():
            t["dram_read_gb"] /= t["launches"]
            t["dram_write_gb"] /= t["launches"]
            t["traffic_gb_per_launch"] = t["dram_read_gb"] + t["dram_write_gb"]
        json.dump({"source": raw, "workload": "bench.py default (cloth 2000 x 128 graphs, pstep 3)",
                   "arithmetic": sys.argv[4] if len(sys.argv) > 4 else "tc", "kernels": traffic},
                  open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()

"""Summarise an `ncu --set full` report (read on the CPU box with `ncu -i ... --page raw --csv`) into profiles/:
a text summary of the headline metrics per captured launch and, for the dominant gemm kernel, the DRAM traffic per launch that
bench.py reports as roofline.traffic."""
import csv
import os
import re
import json
import subprocess
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main(rep, out_txt, title, traffic_json=None, algorithmic=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# {title}", f"# source: {rep} (ncu --set full --clock-control none)"]
    traffic = []
    for row in rows[2:]:
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"{k}: {row[i]} {units[i]}")
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        if re.search(os.environ.get("TRAFFIC_KERNEL_RE", "."), row[hdr.index("Kernel Name")]):   # launches that count for the traffic average
            traffic.append(to_bytes(row[ir], units[ir]) + to_bytes(row[iw], units[iw]))
            traffic_name = row[hdr.index("Kernel Name")]
        lines.append("")
    open(out_txt, "w").write("\n".join(lines) + "\n")
    if traffic_json:
        name = traffic_name
        json.dump(dict(kernel=name.split("(")[0].replace("void <unnamed>::", ""), dram_bytes_per_launch=sum(traffic) / len(traffic),
                       launches_captured=len(traffic), algorithmic_bytes_per_launch=algorithmic, source=out_txt), open(traffic_json, "w"), indent=1)
    print(open(out_txt).read())


if __name__ == "__main__":
    rep, out_txt, title = sys.argv[1:4]
    tj = sys.argv[4] if len(sys.argv) > 4 else None
    alg = float(sys.argv[5]) if len(sys.argv) > 5 else None
    main(rep, out_txt, title, tj, alg)

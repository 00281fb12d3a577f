#!/usr/bin/env python3
"""profiles/gemm_traffic.json (read by bench.py for roofline.traffic) from an ncu --set full summary written by scripts/ncu_summary.py.
usage: scripts/gemm_traffic.py profiles/r02_gemm_ncu_full_summary.txt > profiles/gemm_traffic.json"""
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path):
    launches = []
    for blk in open(path).read().split("---"):
        d = {}
        for ln in blk.splitlines():
            m = re.match(r"\s+(\S.*?) = (.*)", ln)
            if m:
                d[m.group(1)] = m.group(2).strip()
        if "Kernel Name" not in d or "gemm_tc" not in d["Kernel Name"]:
            continue

        def val(key):
            v, u = d[key].split()[:2]
            return float(v.replace(",", "")) * UNIT.get(u, 1.0)
        name = re.search(r"(gemm_tc2?_kernel<[^>]*>)", d["Kernel Name"]).group(1)
        launches.append({"kernel": name, "duration_us": float(d["gpu__time_duration.sum"].split()[0]),
                         "dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                         "tensor_pipe_active_pct": float(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"].split()[0])})
    out = {"source": f"{path} (ncu --set full --clock-control none, {len(launches)} consecutive tcgen05 GEMM launches inside the 1024-stream step; "
                     "dram__bytes_read.sum + dram__bytes_write.sum)",
           "dram_bytes_per_launch": sum(x["dram_bytes"] for x in launches) / max(len(launches), 1), "launches": launches}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])

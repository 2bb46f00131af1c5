#!/usr/bin/env python
"""Digest of an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list
(read on the CPU box):

    python tools/ncu_launches.py gpurun_out/launches.csv [--shares profiles/rXX_launch_shares.txt]
                                                          [--traffic profiles/k1_traffic.json] [--note "..."]

--shares : per-kernel total time, launch count and SHARE of the step (ncu serialises launches and runs them cold-cache:
           the shares are comparable with the CUDA-event step time, the absolutes are not).
--traffic: DRAM bytes (read + written) per K1 launch, averaged over every conv_gemm_tcgen05_kernel launch in the list —
           what bench.py reports as roofline.traffic next to the algorithmic FLOPs per launch.
"""
import csv
import json
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"\(.*$", "", name)  # drop the parameter list
    return name.strip()


def main():
    path = sys.argv[1]
    args = sys.argv[2:]
    opt = {args[i]: args[i + 1] for i in range(0, len(args) - 1, 2) if args[i].startswith("--")}
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ci = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    per_launch = defaultdict(dict)
    names = {}
    for r in rd:
        lid = int(r[ci["ID"]])
        names[lid] = short(r[ci["Kernel Name"]])
        val = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]]
        m = r[ci["Metric Name"]]
        if m == "gpu__time_duration.sum":
            val *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)  # -> us
        elif m.startswith("dram__bytes"):
            val *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        per_launch[lid][m] = val
    tot = defaultdict(float)
    cnt = defaultdict(int)
    dram = defaultdict(float)
    for lid, ms in per_launch.items():
        n = names[lid]
        tot[n] += ms.get("gpu__time_duration.sum", 0.0)
        cnt[n] += 1
        dram[n] += ms.get("dram__bytes_read.sum", 0.0) + ms.get("dram__bytes_write.sum", 0.0)
    total = sum(tot.values())
    out = []
    if "--note" in opt:
        out.append(opt["--note"])
    for n in sorted(tot, key=lambda k: -tot[k]):
        extra = f"  {dram[n] / cnt[n] / 1e6:9.1f} MB DRAM/launch" if dram[n] else ""
        out.append(f"{tot[n]:12.1f} us {cnt[n]:6d} {100 * tot[n] / total:5.1f}%  {n}{extra}")
    out.append(f"{total:12.1f} us total")
    text = "\n".join(out)
    print(text)
    if "--shares" in opt:
        open(opt["--shares"], "w").write(text + "\n")
    if "--traffic" in opt:
        k1 = [n for n in tot if "conv_gemm_tcgen05_kernel" in n]
        nl = sum(cnt[n] for n in k1)
        by = sum(dram[n] for n in k1)
        if nl and by:
            d = {"kernel": "conv_gemm_tcgen05_kernel (all variants of a forward pass)", "launches": nl,
                 "dram_bytes_per_launch": round(by / nl), "dram_bytes_total": round(by),
                 "k1_time_share": round(sum(tot[n] for n in k1) / total, 4),
                 "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum over {path.split('/')[-1]}"
                           + (": " + opt["--note"] if "--note" in opt else "")}
            json.dump(d, open(opt["--traffic"], "w"), indent=1)
            print("wrote", opt["--traffic"], d)


if __name__ == "__main__":
    main()

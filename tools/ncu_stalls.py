#!/usr/bin/env python
"""Warp-stall samples per SASS instruction from an `ncu --set full --import-source on` report (read on the CPU box):

    python tools/ncu_stalls.py gpurun_out/prof.ncu-rep [launch index] [top N]

Prints, per kernel section of the source page, the stall-reason totals and the instructions holding the most samples."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(which), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    name, hdr, data = None, None, []
    for r in rows:
        if r and r[0] == "Kernel Name":
            name = r[1]
        elif r and r[0] == "Address":
            hdr = r
        elif hdr and len(r) == len(hdr):
            data.append(r)
    ci = {k: i for i, k in enumerate(hdr)}
    n = lambda r, k: int(r[ci[k]] or 0)
    tot = sum(n(r, "# Samples") for r in data)
    stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    agg = sorted(((k, sum(n(r, k) for r in data)) for k in stalls), key=lambda kv: -kv[1])
    print(name[:110])
    print(f"{tot} samples over {len(data)} instructions; stall reasons:",
          ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in agg[:8]))
    for r in sorted(data, key=lambda r: -n(r, "# Samples"))[:topn]:
        st = sorted(((k[6:], n(r, k)) for k in stalls), key=lambda kv: -kv[1])[:2]
        print(f'{100 * n(r, "# Samples") / tot:5.1f}%  {r[ci["Address"]][-5:]}  {r[ci["Source"]].strip()[:72]:72s} {st}')


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""profiles/ncu_traffic.json from an ncu capture (not by hand): dram__bytes_read.sum + dram__bytes_write.sum per launch
of the dominant kernel, averaged over the captured launches.
    python tools/ncu_traffic.py gpurun_out/r02f_prof_spmv_insolve.ncu-rep S3-hex-256/1 k_spmv_s3_rt profiles/r02_ncu_insolve_S3hex256_rt.txt"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_bytes(v, unit):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    return float(v.replace(",", "")) * f


def main(rep, key, kernel, summary):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ir, iw, ik, it = (hdr.index(n) for n in ("dram__bytes_read.sum", "dram__bytes_write.sum", "Kernel Name", "gpu__time_duration.sum"))
    sel = [d for d in data if kernel in d[ik]]
    rd = sum(to_bytes(d[ir], units[ir]) for d in sel) / len(sel)
    wr = sum(to_bytes(d[iw], units[iw]) for d in sel) / len(sel)
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    J = json.load(open(path)) if os.path.exists(path) else {}
    J["_comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, written by tools/ncu_traffic.py from "
                     "an `ncu --set full` capture of the kernel as it runs inside the solve; bench.py copies the matching entry into roofline.traffic")
    J[key] = {"kernel": sel[0][ik][:80], "bytes": int(rd + wr), "read": int(rd), "write": int(wr), "launches": len(sel),
              "ms_under_ncu": sum(float(d[it]) for d in sel) / len(sel), "capture": summary}
    json.dump(J, open(path, "w"), indent=1)
    print(key, J[key])


if __name__ == "__main__":
    main(*sys.argv[1:5])

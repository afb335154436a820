#!/usr/bin/env python
"""DRAM traffic of the kernels, as bench.py reports it (`roofline.traffic`).

Two steps, because ncu needs the GPU and the repo's committed evidence lives in profiles/:

  (on the GPU box, under gpurun)
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        --csv --log-file gpurun_out/traffic.csv python tools/profile_trace.py
  (here)
    python tools/ncu_traffic.py gpurun_out/traffic.csv      -> profiles/kernel_traffic.json

The JSON carries the hash of the kernel sources it was captured from (bench.kernel_source_hash);
bench.py reports `traffic: null`, loudly, when the sources have changed since.
Launches of tools/profile_trace.py: trace_kernel #1 = fused trace+splat of 2^24 photons (built-in
scene, 1024^2); splat_kernel, last = 2^25 records; gather_kernel, last = 4096^2 pixels."""
import csv
import datetime
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(open(path)) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r)
    h = rows[hdr]
    kn, mn, mv, mu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    idc = h.index("ID")
    launches = {}
    for r in rows[hdr + 1:]:
        if len(r) != len(h):
            continue
        v = float(r[mv].replace(",", ""))
        unit = r[mu].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
                 "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}.get(unit, 1)
        name = r[kn].split("(")[0].split("<")[0].replace("void ", "").strip()      # "void trace_kernel<1, 0>(...)"
        launches.setdefault((int(r[idc]), name), {})[r[mn]] = v * scale
    by_kernel = {}
    for (i, name), m in sorted(launches.items()):
        by_kernel.setdefault(name, []).append(m)
    def dram(m):
        return m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
    out = {"kernel_source_sha16": bench.kernel_source_hash(),
           "captured": datetime.date.today().isoformat() + " (tools/ncu_traffic.py over tools/profile_trace.py, "
                       "ncu --clock-control none: cold-cache, serialised launches)",
           "kernels": {}}
    tk = by_kernel.get("trace_kernel", [])
    if len(tk) >= 2:
        out["kernels"]["trace_kernel"] = {"photons": 1 << 24, "dram_bytes": dram(tk[1]),
                                          "dram_bytes_per_photon": dram(tk[1]) / (1 << 24),
                                          "seconds": tk[1]["gpu__time_duration.sum"]}
    sk = by_kernel.get("splat_kernel", [])
    if sk:
        out["kernels"]["splat_kernel"] = {"records": 1 << 25, "dram_bytes": dram(sk[-1]),
                                          "seconds": sk[-1]["gpu__time_duration.sum"]}
    gk = by_kernel.get("gather_kernel", [])
    if gk:
        out["kernels"]["gather_kernel"] = {"pixels": 4096 * 4096, "dram_bytes": dram(gk[-1]),
                                           "seconds": gk[-1]["gpu__time_duration.sum"]}
    dst = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full --page raw --csv` export of tools/profile_step.py (one steady-state pass):
dram__bytes_read.sum / dram__bytes_write.sum per launch of the kernels bench.py reports a `traffic` for, stamped with
the sha1 of the GEMM sources so that bench.py drops the figure when the kernel changes.

    python tools/make_traffic.py gpurun_out/r02_full_raw.csv r02
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_sha1  # noqa: E402

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(path, tag):
    r = csv.reader(open(path))
    hdr, units = next(r), next(r)
    idx = {h: i for i, h in enumerate(hdr)}
    out = {"source": "profiles/%s_full_summary.md (ncu --set full --clock-control none, one steady-state pass, B200)" % tag,
           "gemm_source_sha1": kernel_source_sha1()}
    seen = {}
    for row in r:
        name = re.sub(r"\(.*", "", row[idx["Kernel Name"]]).replace("void ", "").replace("gk::", "")
        vals = {}
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = idx[key]
            vals[key] = float(row[i].replace(",", "")) * UNIT.get(units[i], 1.0)
        seen.setdefault(name, []).append({"dram_read_bytes": vals["dram__bytes_read.sum"], "dram_write_bytes": vals["dram__bytes_write.sum"]})
    g = seen.get("gram_tc2_kernel", [])
    if len(g) >= 1:
        out["gram_tc2_kernel hybrid (head columns)"] = g[0]
    if len(g) >= 2:
        out["gram_tc2_kernel dense (all shared columns)"] = g[1]
    for k in ("tail_pairs<float>", "wl_fused2_kernel"):
        for name in seen:  # template instances print as wl_fused2_kernel<1> / <true>
            if name == k or name.startswith(k + "<"):
                out[k] = seen[name][0]
                break
    json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "rXX")

#!/usr/bin/env python
"""Sum DRAM traffic and duration per kernel from an
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` launch list of ONE
train step (tools/profile_step.py) and write the JSON bench.py reports as roofline.traffic."""
import collections
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0,
        "ms": 1e3, "msecond": 1e3}


def main(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: {"launches": set(), "dram_bytes": 0.0, "us": 0.0})
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", "")) * UNIT.get(row["Metric Unit"], 1.0)
        a = agg[name]
        a["launches"].add(row["ID"])
        if row["Metric Name"].startswith("dram__bytes"):
            a["dram_bytes"] += v
        elif row["Metric Name"].startswith("gpu__time_duration"):
            a["us"] += v
    res = {k: {"launches": len(v["launches"]), "dram_bytes_per_step": v["dram_bytes"], "us_per_step": v["us"]}
           for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["us"])}
    tc = [v for k, v in res.items() if "tapconv_tc_kernel" in k or "tapwgrad_tc_kernel" in k or "resunit_tc_kernel" in k]
    summary = {"source": path, "tensor_core_kernels": {"launches": sum(v["launches"] for v in tc),
                                                        "dram_bytes_per_step": sum(v["dram_bytes_per_step"] for v in tc),
                                                        "us_per_step_serialised": sum(v["us_per_step"] for v in tc)},
               "all_kernels_dram_bytes_per_step": sum(v["dram_bytes_per_step"] for v in res.values()),
               "per_kernel": res}
    json.dump(summary, open(out, "w"), indent=1)
    print(json.dumps(summary["tensor_core_kernels"]), summary["all_kernels_dram_bytes_per_step"])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

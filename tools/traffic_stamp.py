#!/usr/bin/env python
"""Write profiles-style roofline_traffic.json from ncu summary CSVs (tools/ncu_summary.py output).

    python tools/traffic_stamp.py <tag> <out.json> <workload:kind:kernel-regex:summary.csv> ...

Each entry records dram__bytes_read.sum + dram__bytes_write.sum of the first launch whose name
matches the regex, the capture's tag and the hash of the library sources it was taken on
(bench.py: src_sha16), so that bench.py can tell whether the number belongs to the binary it times."""
import csv
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import src_sha16  # noqa: E402

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    tag, out = sys.argv[1], sys.argv[2]
    res = {"_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none "
                      "(tools/gpu_round.sh); src_sha16 = bench.py:src_sha16() at capture time"}
    for spec in sys.argv[3:]:
        wl, kind, rx, path = spec.split(":", 3)
        if not os.path.exists(path):
            continue
        tot, kname = 0.0, None
        for r in csv.reader(open(path)):
            if len(r) < 4 or not re.search(rx, r[0]):
                continue
            if kname is None:
                kname = r[0]
            if r[0] != kname:
                continue
            if r[1] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(r[2]) * UNIT.get(r[3], 1)
        if kname:
            res[f"{wl}:{kind}"] = {"bytes": int(tot), "tag": tag, "kernel": kname, "src_sha16": src_sha16()}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""tests/golden/job_light.json.gz: the 70 job-light join queries and their true cardinalities, copied from the
reference's workload file Benchmark/IMDB/job-light.sql (data, "<sql>||<true cardinality>" per line), so that the GPU box
(no /root/reference) can run the planner tests.      python tools/make_golden_joblight.py [/root/reference]"""
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    rows = []
    with open(os.path.join(ref, "Benchmark", "IMDB", "job-light.sql")) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            sql, true = line.rsplit("||", 1)
            rows.append({"sql": sql.strip(), "true": int(true)})
    out = os.path.join(ROOT, "tests", "golden", "job_light.json.gz")
    with gzip.open(out, "wt") as g:
        json.dump({"source": "Benchmark/IMDB/job-light.sql", "paper_table9_qerror_50_90_95_100": [1.30, 3.534, 4.836, 19.13],
                   "queries": rows}, g)
    print(out, len(rows))


if __name__ == "__main__":
    main()

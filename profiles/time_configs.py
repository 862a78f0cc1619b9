#!/usr/bin/env python
"""Per-config timing table: a thin front end of bench_configs.py (the measurement
bench.py embeds as `configs`), kept under this name for the ncu recipes in profiles/.
    python profiles/time_configs.py 5a        -> one JSON line, this process
    python profiles/time_configs.py 1 2 3     -> one fresh process per config"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_configs  # noqa: E402

if __name__ == "__main__":
    which = sys.argv[1:] or bench_configs.ALL
    if len(which) == 1:
        print(json.dumps(bench_configs.run_one(which[0])), flush=True)
    else:
        for rec in bench_configs.run_all(which):
            print(json.dumps(rec), flush=True)

"""Condense the tolerance ledger (gpurun_out/parity_errors.json, written by the -m gpu suite) into one row per (test, quantity):
python tools/parity_summary.py gpurun_out/parity_errors.json > profiles/r2_parity_summary.csv"""
import collections
import csv
import json
import re
import sys

entries = json.load(open(sys.argv[1]))
groups = collections.OrderedDict()
for e in entries:
    test = re.sub(r"\[.*\]$", "", e["test"].split("::")[-1])
    groups.setdefault((e["kind"], test, e["what"]), []).append(e)
w = csv.writer(sys.stdout)
w.writerow(["kind", "test", "quantity", "comparisons", "max_abs_err", "max_abs_err_over_tensor_max", "max_rel_err_entries_above_1e-3_of_max",
            "rtol", "atol_last", "min_headroom"])
for (kind, test, what), es in groups.items():
    w.writerow([kind, test, what, len(es), "%.3e" % max(e["max_abs_err"] for e in es), "%.3e" % max(e["max_abs_err_over_scale"] for e in es),
                "%.3e" % max(e["max_rel_err_big_entries"] for e in es), es[-1]["rtol"], "%.3g" % es[-1]["atol"],
                "%.1f" % min(min(e["headroom"], 1e9) for e in es)])

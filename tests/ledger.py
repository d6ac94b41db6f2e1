"""Tolerance ledger: every parity comparison of the GPU suite records its measured error next to the tolerance it was
checked against; the session writes them to gpurun_out/parity_errors.json (copied to profiles/ per round)."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENTRIES = []


def _current_test():
    return os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0]


def record(what, a, b, rtol, atol, kind="cuda_vs_oracle"):
    """a = value under test, b = reference (numpy, any shape).  max_rel_err is taken over the entries that matter: |b| above
    1e-3 of the tensor's largest magnitude."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.size == 0:
        return
    diff = np.abs(a - b)
    scale = float(np.abs(b).max())
    big = np.abs(b) > 1e-3 * scale if scale > 0 else np.zeros_like(b, bool)
    ENTRIES.append({"test": _current_test(), "what": what, "kind": kind, "n": int(a.size), "ref_max_abs": scale,
                    "max_abs_err": float(diff.max()), "max_abs_err_over_scale": float(diff.max() / scale) if scale > 0 else 0.0,
                    "max_rel_err_big_entries": float((diff[big] / np.abs(b[big])).max()) if big.any() else 0.0,
                    "rtol": rtol, "atol": atol,
                    "headroom": float(((atol + rtol * np.abs(b)) / np.maximum(diff, 1e-300)).min())})


def dump():
    if not ENTRIES:
        return
    path = os.path.join(ROOT, "gpurun_out", "parity_errors.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as fh:
        json.dump(ENTRIES, fh, indent=1)

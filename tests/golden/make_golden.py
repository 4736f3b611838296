"""Writes the golden fixtures of tests/golden/: seeded inputs and the outputs of the CPU oracle for them.

The reference is Rust and cannot be built in this image (no cargo / rustc), so the vectors cannot come from
the reference itself; they come from oracle/mdb_oracle.cc, which tests/test_oracle_golden.py pins against every
golden vector and known-answer test the reference holds for this path.  The fixtures freeze those outputs:
`pytest -m "not gpu"` checks that the oracle still reproduces them, `pytest -m gpu` that the CUDA path does.

Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import mdb_oracle as O  # noqa: E402
from tests.parity_cases import small_cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SEGMENT_COLUMNS = ("model_type_id", "start_time", "end_time", "min_value", "max_value", "timestamps_off", "timestamps_data",
                   "values_off", "values_data", "residuals_off", "residuals_data", "unit_seg_off")


def selected_cases():
    keep = ("mixed-irr0-noiseFalse-eb(2, 5.0)", "mixed-irr1-noiseTrue-eb(0, 0.0)", "mixed-irr1-noiseTrue-eb(1, 5.0)",
            "sine-epoch-eb(2, 1.0)", "walk-epoch-eb(0, 0.0)", "walk-irregular-rel1", "ragged-units", "specials-eb(0, 0.0)",
            "specials-eb(2, 10.0)", "special-runs-eb(1, 1.0)", "long-residual-runs-eb(2, 1.0)", "lossy-extreme-eb(2, 1e-06)",
            "lossy-extreme-eb(1, 3e+38)")
    cases = {c[0]: c for c in small_cases()}
    return [cases[k] for k in keep]


def file_name(case_name):
    return "".join(ch if ch.isalnum() else "_" for ch in case_name).strip("_") + ".npz"


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def main():
    for name, ts, vals, off, ebs in selected_cases():
        seg = O.compress(ts, vals, off, eb=ebs)
        gts, gval, _ = O.grid(seg)
        count, mn, mx, sm = O.aggregate(seg, seg.unit_seg_off)
        out = {"in_timestamps": ts, "in_values": vals, "in_unit_off": off, "in_eb_kind": np.array([e[0] for e in ebs], np.uint8),
               "in_eb_value": np.array([e[1] for e in ebs], np.float32), "grid_timestamps_sha256": sha(gts), "grid_values_sha256": sha(gval),
               "grid_points": np.array([len(gts)], np.uint64), "agg_count": count, "agg_min": mn, "agg_max": mx, "agg_sum": sm}
        for c in SEGMENT_COLUMNS:
            out["seg_" + c] = getattr(seg, c)
        path = os.path.join(HERE, file_name(name))
        np.savez_compressed(path, **out)
        print(f"{os.path.basename(path):60s} {len(ts):7d} points {len(seg):6d} rows {os.path.getsize(path):8d} bytes")


if __name__ == "__main__":
    main()

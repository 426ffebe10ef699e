"""RTTM golden vectors from the REAL reference function (FS-EEND/train/utils/make_rttm.py:10-28, read-only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_rttm.py

The reference module imports h5py at the top (not installed here, unused by make_rttm): a stub module is registered
before the import so that the unmodified function runs.
"""
import json
import os
import sys
import types

import numpy as np
import torch

sys.dont_write_bytecode = True
sys.modules.setdefault("h5py", types.ModuleType("h5py"))
sys.path.insert(0, "/root/reference/FS-EEND")
from train.utils.make_rttm import make_rttm  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.postproc_oracle import synthetic_posteriors  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


CASES = {"rttm_T500_C4": (500, 4, 0, 11, 0.5), "rttm_T37_C3_med5": (37, 3, 1, 5, 0.4), "rttm_T1200_C8": (1200, 8, 2, 11, 0.5),
         "rttm_T9_C2": (9, 2, 3, 11, 0.5), "rttm_T300_C5_nomed": (300, 5, 4, 1, 0.5)}


def main():
    out = {}
    for name, (T, C, seed, median, thr) in CASES.items():
        pred = synthetic_posteriors(T, C, seed)
        rttm = make_rttm("rec_" + name, pred, frame_shift=80, threshold=thr, median=median, subsampling=10, sampling_rate=8000)
        out[name] = {"T": T, "C": C, "seed": seed, "median": median, "threshold": thr,
                     "rttm": {k: v for k, v in rttm.items()}, "n_lines": sum(len(v) for v in rttm.values())}
        print(name, "lines", out[name]["n_lines"])
    with open(os.path.join(HERE, "rttm_golden.json"), "w") as f:
        json.dump(out, f, indent=0)


if __name__ == "__main__":
    main()

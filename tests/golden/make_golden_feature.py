"""Golden vectors for splice + subsample from the REAL reference functions (FS-EEND/datasets/feature.py:103-133).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_feature.py

librosa and soundfile (absent here; unused by splice / subsample) are stubbed before the import.
"""
import os
import sys
import types

import numpy as np

sys.dont_write_bytecode = True
for name in ("librosa", "soundfile"):
    sys.modules.setdefault(name, types.ModuleType(name))
import importlib.util  # noqa: E402

# loaded by file path: the reference's datasets/ has no __init__.py and would lose against an installed `datasets` package
_spec = importlib.util.spec_from_file_location("ref_feature", "/root/reference/FS-EEND/datasets/feature.py")
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
splice, subsample = _mod.splice, _mod.subsample

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"feat_T1003_F23": (1003, 23, 7, 10, 0), "feat_T40_F5_c2_s3": (40, 5, 2, 3, 1), "feat_T7_F23": (7, 23, 7, 10, 2),
         "feat_T500_F23_s1": (500, 23, 7, 1, 3)}


def main():
    rec = {}
    for name, (T, F, ctx, sub, seed) in CASES.items():
        y = np.random.default_rng(seed).standard_normal((T, F)).astype(np.float32)
        ys, _ = subsample(np.ascontiguousarray(splice(y, ctx)), np.zeros((T, 1)), sub)
        rec[name] = np.ascontiguousarray(ys)
        print(name, ys.shape)
    np.savez_compressed(os.path.join(HERE, "feature_golden.npz"), **rec)


if __name__ == "__main__":
    main()

"""Fixture generator (build container only): dumps the UNMODIFIED reference's make_cfg(dataset) for every dataset name
into tests/golden/reference_configs.json (paths as strings).  `easydict` (absent offline) is replaced by this repo's
own EasyDict work-alike for the import.
    python tests/tools/gen_config_golden.py"""
import json
import os
import sys
import types
from pathlib import Path

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bufferx_b200  # noqa: E402
from bufferx_b200.easydict import EasyDict  # noqa: E402

stub = types.ModuleType("easydict")
stub.EasyDict = EasyDict
sys.modules["easydict"] = stub
sys.path.insert(0, "/root/reference")
sys.modules.pop("config", None)
import config as ref_config  # noqa: E402  (the reference's package)

NAMES = ["3DMatch", "3DLoMatch", "Scannetpp_iphone", "Scannetpp_faro", "TIERS", "TIERS_hetero", "KITTI", "WOD", "MIT", "KAIST",
         "KAIST_hetero", "ETH", "Oxford", "ModelNet40"]


def plain(x):
    if isinstance(x, dict):
        return {k: plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [plain(v) for v in x]
    if isinstance(x, Path):
        return str(x)
    if isinstance(x, (int, float, str, bool)) or x is None:
        return x
    return repr(x)


out = {n: plain(ref_config.make_cfg(n, "../datasets")) for n in NAMES}
p = os.path.join(ROOT, "tests", "golden", "reference_configs.json")
json.dump(out, open(p, "w"), indent=1, sort_keys=True)
print("wrote", p, os.path.getsize(p), "bytes;", {n: len(json.dumps(v)) for n, v in out.items()})

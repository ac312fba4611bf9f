"""A/B plugin variants of the C2 Brownian model for tools/ab_plugins.py (built here, run on the GPU box).

    python tools/build_ab_variants.py   ->  Brownian[_16M]_<variant> for every entry of VARIANTS
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ev_diffusion_b200 import build as B, workloads as W

VARIANTS = {
    "ifc0": dict(if_convert=0),      # the compiler's own if-conversion budget (35 SASS instructions per pair of messages on C2)
    "ifc16": dict(if_convert=16),    # the default build (27)
}

if __name__ == "__main__":
    ff = W.brownian_function_files()
    for key in ("c2", "c2_16m"):
        cfg = W.BROWNIAN[key]
        xml = W.brownian_xml(cfg)
        for name, kw in VARIANTS.items():
            if len(sys.argv) > 1 and name not in sys.argv[1:]:
                continue
            print(B.build_model(xml, cfg.name + "_" + name, fmad=True, function_files=ff, **kw))

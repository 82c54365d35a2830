"""A/B builds of kernel variants: lib/variants/libf1l_<name>.so, selected at run time with
F1L_LIB=<path>.

    python tools/build_variants.py name:DEF=VAL,DEF=VAL ...
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from f1tenth_planning_b200 import build as B  # noqa: E402

vdir = os.path.join(B.OUT_DIR, "variants")
os.makedirs(vdir, exist_ok=True)
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    out = os.path.join(vdir, "libf1l_%s.so" % name)
    B.build(out=out, defs=[d for d in defs.split(",") if d], verbose=True)
    print(out)

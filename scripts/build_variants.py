#!/usr/bin/env python
"""Experimental builds of the library next to the product one (A/B runs with PB200_LIB=...): name=DEF1,DEF2 ...
PB200_VARIANT_UNITS=k_s2,k_s3 recompiles only those units (the rest are the product objects)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from posidonius_b200 import build as b  # noqa: E402

for spec in sys.argv[1:]:
    name, _, defs = spec.partition("=")
    out = os.path.join(ROOT, "posidonius_b200", "libpb200_%s.so" % name)
    only = set(os.environ["PB200_VARIANT_UNITS"].split(",")) if os.environ.get("PB200_VARIANT_UNITS") else None
    b.build(force=True, out=out, defines=[d for d in defs.split(",") if d], only=only)
    print(out)

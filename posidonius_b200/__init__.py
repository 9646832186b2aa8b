"""posidonius_b200 — B200-native ensemble implementation of the Posidonius WHFast integration step.

The product is the CUDA library `libposidonius_b200.so` behind the C ABI of include/posidonius_b200.h;
this package is the thin host-side mirror of the reference's Integrator interface around it.
"""
from . import abi  # noqa: F401
from .case import (CaseTables, InvalidCaseError, UnsupportedCaseError, case_from_dict, copy_case,  # noqa: F401
                   load_case_json)

__all__ = ["abi", "CaseTables", "InvalidCaseError", "UnsupportedCaseError", "case_from_dict", "copy_case", "load_case_json"]

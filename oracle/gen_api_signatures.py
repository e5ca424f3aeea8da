#!/usr/bin/env python
"""Record the public API surface of the UNMODIFIED reference's src/lib (TEST INFRASTRUCTURE ONLY).

    python oracle/gen_api_signatures.py [--ref /root/reference] [--out tests/golden/api_signatures.json]

Writes, for src/lib/{SolutionsManagers,ReducedBasis,Estimators}.py, every module-level function, class, method
(parameter names, kinds and defaults) and plain constant.  tests/test_api_signatures.py compares the drop-in mirror
(romhighcontrast_b200/lib) with this file: SURVEY 8b "signatures to keep".  Runs only where /root/reference exists; the
JSON travels with the repo."""
from __future__ import annotations

import argparse
import inspect
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden import import_reference          # noqa: E402


def _default(v):
    if v is inspect.Parameter.empty:
        return None
    if isinstance(v, (int, float, str, bool, type(None))):
        return {"value": v}
    if isinstance(v, (tuple, list)):
        return {"repr": repr(v)}
    return {"repr": type(v).__name__}


def _sig(fn):
    try:
        sig = inspect.signature(fn)
    except (TypeError, ValueError):
        return None
    return [{"name": p.name, "kind": p.kind.name, "default": _default(p.default)} for p in sig.parameters.values()]


def describe(mod):
    out = {"functions": {}, "classes": {}, "constants": {}}
    for name, obj in vars(mod).items():
        if name.startswith("_"):
            continue
        if inspect.isfunction(obj) and obj.__module__ == mod.__name__:
            out["functions"][name] = _sig(obj)
        elif inspect.isclass(obj) and obj.__module__ == mod.__name__:
            members = {}
            for mname, m in vars(obj).items():
                if mname.startswith("__") and mname not in ("__init__", "__getitem__", "__str__"):
                    continue
                raw = m.__func__ if isinstance(m, (staticmethod, classmethod)) else m
                if inspect.isfunction(raw):
                    members[mname] = {"kind": type(m).__name__ if isinstance(m, (staticmethod, classmethod)) else "method",
                                      "params": _sig(raw)}
                elif isinstance(m, property):
                    members[mname] = {"kind": "property", "params": None}
            out["classes"][name] = {"bases": [b.__name__ for b in obj.__bases__], "members": members}
        elif isinstance(obj, (int, float, str)) and not inspect.ismodule(obj):
            out["constants"][name] = obj
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                  "tests", "golden", "api_signatures.json"))
    args = ap.parse_args()
    SM, RB, ES = import_reference(args.ref)
    api = {"SolutionsManagers": describe(SM), "ReducedBasis": describe(RB), "Estimators": describe(ES)}
    # instance attributes right after construction with the constructors' defaults / the experiment driver's arguments
    # (plotting code reads .name / .linestyle / .greedy_for / .add_inf_solutions, HighContrast.py:45-56,236-241)
    def attrs(obj):
        return {k: ({"value": v} if isinstance(v, (int, float, str, bool, type(None))) else {"type": type(v).__name__})
                for k, v in sorted(vars(obj).items())}
    api["instances"] = {
        "ReducedBasisGreedy()": attrs(RB.ReducedBasisGreedy()),
        "ReducedBasisGreedy(greedy_for=GREEDY_FOR_H10)": attrs(RB.ReducedBasisGreedy(greedy_for=RB.GREEDY_FOR_H10)),
        "ReducedBasisRandom()": attrs(RB.ReducedBasisRandom()),
        "ReducedBasisRandom(False)": attrs(RB.ReducedBasisRandom(False)),
        "ReducedBasisPCA()": attrs(RB.ReducedBasisPCA()),
        "ReducedBasisPCA(False)": attrs(RB.ReducedBasisPCA(False)),
        "BaseReducedBasis()": attrs(RB.BaseReducedBasis()),
        "SolutionsManagerFEM((2, 2), 3)": attrs(SM.SolutionsManagerFEM((2, 2), 3)),
    }
    with open(args.out, "w") as f:
        json.dump(api, f, indent=1, sort_keys=True)
    n = sum(len(m["functions"]) + sum(len(c["members"]) for c in m["classes"].values())
            for k, m in api.items() if k != "instances")
    print(f"wrote {args.out}: {n} callables")


if __name__ == "__main__":
    main()

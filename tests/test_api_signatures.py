"""Drop-in boundary ("signatures to keep"): the mirror of src/lib keeps every in-scope name of the reference
with the same parameters, order, kinds and defaults; additions are allowed only as trailing parameters with defaults.

tests/golden/api_signatures.json was written by oracle/gen_api_signatures.py from the unmodified reference."""
import importlib
import inspect
import json
import os

import pytest

from conftest import GOLDEN

API = json.load(open(os.path.join(GOLDEN, "api_signatures.json")))

# out of scope: never instantiated by the reference, or not constructible in it
OUT_OF_SCOPE = {("SolutionsManagers", "SolutionsManagerPolynomial"), ("SolutionsManagers", "init_polynomial_variables"),
                ("SolutionsManagers", "h1_error"),
                ("Estimators", "EstimatorNN"), ("Estimators", "EstimatorNear"), ("Estimators", "EstimatorTree")}


def _default(v):
    if v is inspect.Parameter.empty:
        return None
    if isinstance(v, (int, float, str, bool, type(None))):
        return {"value": v}
    if isinstance(v, (tuple, list)):
        return {"repr": repr(v)}
    return {"repr": type(v).__name__}


def _check_params(where, ref_params, fn):
    got = list(inspect.signature(fn).parameters.values())
    assert len(got) >= len([p for p in ref_params if p["kind"] != "VAR_KEYWORD"]), where
    ref_named = [p for p in ref_params if p["kind"] not in ("VAR_KEYWORD", "VAR_POSITIONAL")]
    for i, rp in enumerate(ref_named):
        gp = got[i]
        assert gp.name == rp["name"], f"{where}: parameter {i} is {gp.name!r}, reference has {rp['name']!r}"
        assert gp.kind.name == rp["kind"], f"{where}: kind of {gp.name}"
        assert _default(gp.default) == rp["default"], f"{where}: default of {gp.name}: {gp.default!r} vs {rp['default']}"
    extra = [p for p in got[len(ref_named):] if p.kind.name not in ("VAR_KEYWORD", "VAR_POSITIONAL")]
    for p in extra:                                            # additive parameters must not change existing calls
        assert p.default is not inspect.Parameter.empty, f"{where}: added parameter {p.name} has no default"
    if any(p["kind"] == "VAR_KEYWORD" for p in ref_params):    # build(**kwargs) swallows unknown keywords: keep that
        assert any(p.kind.name == "VAR_KEYWORD" for p in got), f"{where}: reference accepts **kwargs"


@pytest.mark.parametrize("spelling", ["romhighcontrast_b200.lib", "src.lib", "lib"])
@pytest.mark.parametrize("modname", sorted(k for k in API if k != "instances"))
def test_mirror_keeps_the_reference_signatures(spelling, modname):
    mod = importlib.import_module(f"{spelling}.{modname}")
    ref = API[modname]
    for name, value in ref["constants"].items():
        assert getattr(mod, name) == value, name
    for name, params in ref["functions"].items():
        if (modname, name) in OUT_OF_SCOPE:
            continue
        _check_params(f"{modname}.{name}", params, getattr(mod, name))
    for cname, cdesc in ref["classes"].items():
        if (modname, cname) in OUT_OF_SCOPE:
            continue
        cls = getattr(mod, cname)
        assert [b.__name__ for b in cls.__bases__] == cdesc["bases"], cname
        for mname, m in cdesc["members"].items():
            where = f"{modname}.{cname}.{mname}"
            assert hasattr(cls, mname), where
            attr = inspect.getattr_static(cls, mname)
            if m["kind"] == "property":
                assert isinstance(attr, property), where
                continue
            if m["kind"] in ("staticmethod", "classmethod"):
                assert type(attr).__name__ == m["kind"], where
            _check_params(where, m["params"], getattr(cls, mname) if m["kind"] != "method" else attr)


def test_instances_expose_the_reference_attributes():
    """Every attribute a freshly constructed reference object has (plotting code reads .name / .linestyle / .greedy_for /
    .add_inf_solutions, HighContrast.py:45-56,236-241) exists on the mirror's object with the same simple value or the same
    container type.  The dense A_preassembled tensors are lazy properties here and are only checked for presence."""
    import lib.ReducedBasis as RB
    import lib.SolutionsManagers as SM
    make = {
        "ReducedBasisGreedy()": lambda: RB.ReducedBasisGreedy(),
        "ReducedBasisGreedy(greedy_for=GREEDY_FOR_H10)": lambda: RB.ReducedBasisGreedy(greedy_for=RB.GREEDY_FOR_H10),
        "ReducedBasisRandom()": lambda: RB.ReducedBasisRandom(),
        "ReducedBasisRandom(False)": lambda: RB.ReducedBasisRandom(False),
        "ReducedBasisPCA()": lambda: RB.ReducedBasisPCA(),
        "ReducedBasisPCA(False)": lambda: RB.ReducedBasisPCA(False),
        "BaseReducedBasis()": lambda: RB.BaseReducedBasis(),
        "SolutionsManagerFEM((2, 2), 3)": lambda: SM.SolutionsManagerFEM((2, 2), 3),
    }
    assert set(make) == set(API["instances"])
    lazy = {"A_preassembled", "A_preassembled4h1_norm"}
    for key, attrs in API["instances"].items():
        obj = make[key]()
        for name, desc in attrs.items():
            where = f"{key}.{name}"
            if name in lazy:
                assert isinstance(inspect.getattr_static(type(obj), name), property), where
                continue
            assert hasattr(obj, name), where
            value = getattr(obj, name)
            if "value" in desc:
                assert value == desc["value"] and type(value) is type(desc["value"]), f"{where}: {value!r} vs {desc['value']!r}"
            else:
                assert type(value).__name__ == desc["type"], f"{where}: {type(value).__name__} vs {desc['type']}"

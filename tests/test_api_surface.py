"""Drop-in boundary: every public function of the reference's path modules and every method of its ``NDMPS``
exists in this package under the same import path, with the same positional parameters in the same order and the same
defaults; what the package adds is keyword-only.  The reference side is ``tests/golden/reference_api.json``, read from
the reference's source with ``ast`` (``tests/golden/make_golden_api.py``)."""
import importlib
import inspect
import json

import numpy as np
import pytest

from conftest import GOLDEN

API = json.loads((GOLDEN / "reference_api.json").read_text())


def _default(src):
    return eval(src, {"np": np})          # the reference's defaults are literals and np.uint8 / np.uint16


def _check(ours, want, where):
    sig = inspect.signature(ours)
    positional = [p for p in sig.parameters.values() if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
    names = [p.name for p in positional]
    ref_names = list(want["params"])
    if ref_names and ref_names[0] in ("self", "cls") and (not names or names[0] != ref_names[0]):
        ref_names, ref_defaults = ref_names[1:], want["defaults"][1:]          # bound / classmethod view
    else:
        ref_defaults = want["defaults"]
    assert names == ref_names, f"{where}: positional parameters {names} != reference {ref_names}"
    for p, d in zip(positional, ref_defaults):
        if d is None:
            assert p.default is inspect.Parameter.empty, f"{where}: {p.name} has a default the reference does not have"
        else:
            assert p.default is not inspect.Parameter.empty, f"{where}: {p.name} lost its default {d}"
            got, ref = p.default, _default(d)
            assert got is ref or got == ref, f"{where}: default of {p.name} is {got!r}, reference {ref!r}"
    extra = [p.name for p in sig.parameters.values() if p.kind == p.KEYWORD_ONLY and p.name not in want["kwonly"]]
    assert all(sig.parameters[n].default is not inspect.Parameter.empty for n in extra), f"{where}: required keyword-only extras {extra}"


@pytest.mark.parametrize("module", sorted(API))
def test_module_functions_match_the_reference(module):
    ours = importlib.import_module("imgcompressionmps." + module)
    for name, want in API[module]["functions"].items():
        assert hasattr(ours, name), f"imgcompressionmps.{module}.{name} is missing"
        _check(getattr(ours, name), want, f"{module}.{name}")


def test_ndmps_methods_match_the_reference():
    from imgcompressionmps.core.ndmps import NDMPS
    methods = API["core.ndmps"]["classes"]["NDMPS"]
    assert len(methods) == 17
    for name, want in methods.items():
        assert hasattr(NDMPS, name), f"NDMPS.{name} is missing"
        raw = inspect.getattr_static(NDMPS, name)
        assert isinstance(raw, classmethod) == ("classmethod" in want["decorators"]), f"NDMPS.{name}: classmethod-ness differs"
        fn = raw.__func__ if isinstance(raw, classmethod) else raw
        _check(fn, want, f"NDMPS.{name}")

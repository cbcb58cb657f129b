"""Drop-in check: every public function of the reference modules this package mirrors exists here under the same
module path and name, with the same positional parameters in the same order and the same defaults
(tests/golden/signatures.json is written from the real reference by oracle/gen_signatures.py).  The NumPy module and
its torch twin are ONE implementation here, importable under both module paths (`quat` and `quat_torch`, ...);
where the twins name a parameter differently (`axis` / `dim`) both are accepted.  CPU only: signatures are inspected, nothing is launched."""
import importlib
import inspect
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
TABLE = json.load(open(os.path.join(HERE, "golden", "signatures.json")))
# parameters the two twins of the reference name differently; this package accepts either spelling
ALIASES = {"axis": {"axis", "dim"}, "dim": {"axis", "dim"}}


@pytest.mark.parametrize("ref_module", sorted(TABLE))
def test_public_surface_and_signatures(ref_module):
    ours = importlib.import_module("pymotion_b200." + ref_module)  # the twin's own module path (alias modules)
    for name, params in TABLE[ref_module].items():
        fn = getattr(ours, name, None)
        assert callable(fn), f"{ref_module}.{name} is missing"
        mine = list(inspect.signature(fn).parameters.items())
        names = [n for n, _ in mine]
        for pos, (pname, pdefault) in enumerate(params):
            accepted = ALIASES.get(pname, {pname})
            if pos < len(names) and names[pos] in accepted:
                got = mine[pos][1]
            else:  # an alias may sit further back as a keyword (axis=None, dim=None)
                hits = [p for n, p in mine if n in accepted]
                assert hits and pname in names, f"{ref_module}.{name}: parameter {pname!r} missing (have {names})"
                got = dict(mine)[pname]
            if pdefault is not None and pname not in ALIASES:
                assert got.default is not inspect.Parameter.empty, f"{ref_module}.{name}({pname}) lost its default"
                assert repr(got.default) == pdefault, f"{ref_module}.{name}({pname}) default {got.default!r} != {pdefault}"
        required = [n for n, p in mine if p.default is inspect.Parameter.empty]
        ref_required = [p for p, d in params if d is None]
        assert len(required) <= len(ref_required), f"{ref_module}.{name} needs more arguments than the reference: {required}"

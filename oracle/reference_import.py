"""ORACLE support (test infrastructure): import the REAL reference modules src/utils.py,
src/models.py and src/predictor.py, unmodified, from /root/reference.

They only need torch plus `peft.get_peft_model_state_dict` (imported at utils.py:18 and used only by
the kohya export helper), so a 2-line stub `peft` module is placed in sys.modules for the import.
/root/reference does not exist on the GPU box: callers must check `available()` and the GPU tests /
smoke / bench never call this.  Used by tests/test_oracle_glue.py and
tests/golden/make_golden.py (which writes the committed golden vectors).
"""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path("/root/reference")


def available() -> bool:
    return (REFERENCE_ROOT / "src" / "predictor.py").exists()


def load():
    """Returns (utils, models, predictor) modules of the reference."""
    if not available():
        raise RuntimeError("/root/reference is not present on this machine")
    if "peft" not in sys.modules:
        stub = types.ModuleType("peft")
        stub.get_peft_model_state_dict = lambda *a, **k: {}
        sys.modules["peft"] = stub
        stubbed = True
    else:
        stubbed = False
    # the reference is a package called `src`; import it under an alias so it cannot shadow anything
    spec = importlib.util.spec_from_file_location(
        "tdm_reference_src", REFERENCE_ROOT / "src" / "__init__.py",
        submodule_search_locations=[str(REFERENCE_ROOT / "src")])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules["tdm_reference_src"] = pkg
    spec.loader.exec_module(pkg)
    utils = importlib.import_module("tdm_reference_src.utils")
    models = importlib.import_module("tdm_reference_src.models")
    predictor = importlib.import_module("tdm_reference_src.predictor")
    if stubbed:
        del sys.modules["peft"]
    return utils, models, predictor

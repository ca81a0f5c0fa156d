"""Import shim: `import medvill_b200` loads the package that lives in `multi-modality-self-supervision_b200/`
(the directory name the project layout prescribes is not a valid Python identifier)."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multi-modality-self-supervision_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)

"""Import alias: `import tron_b200` loads the package stored in `deep-q-learning_tron_b200/`.

The directory name is fixed by the project layout and is not a valid Python identifier, so this
10-line loader registers it under the importable name `tron_b200`.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "deep-q-learning_tron_b200")
_spec = _ilu.spec_from_file_location("tron_b200", _os.path.join(_dir, "__init__.py"),
                                     submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["tron_b200"] = _mod
_spec.loader.exec_module(_mod)

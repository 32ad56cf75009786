"""Importable alias of the package directory ``multimodal-image-similarity-search_b200/``
(hyphens make the directory name itself un-importable).  ``import mmiss_b200`` yields the real
package, with submodules (``mmiss_b200.collection`` ...) resolving inside that directory."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "multimodal-image-similarity-search_b200")
_spec = _u.spec_from_file_location("mmiss_b200", _os.path.join(_dir, "__init__.py"),
                                   submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["mmiss_b200"] = _mod
_spec.loader.exec_module(_mod)

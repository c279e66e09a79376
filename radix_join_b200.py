"""Import alias: the package directory is `radix-join_b200/` (the name the build contract fixes),
which is not a Python identifier.  `import radix_join_b200` loads that directory as a package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "radix-join_b200")
_spec = importlib.util.spec_from_file_location(
    "radix_join_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["radix_join_b200"] = _mod
_spec.loader.exec_module(_mod)

"""Import alias: `import spark_agd_b200` loads the package that lives in `spark-agd_b200/`
(a hyphen is not a valid Python identifier, so the directory cannot be imported by name)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "spark-agd_b200")
_spec = importlib.util.spec_from_file_location("spark_agd_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["spark_agd_b200"] = _mod
_spec.loader.exec_module(_mod)

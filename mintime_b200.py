"""Import alias: ``import mintime_b200`` loads the package directory
``mintime-multi-identity-size-invariant-timesformer-for-video-deepfake-detection_b200/`` (whose
name is not a Python identifier) under the module name ``mintime_b200``."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "mintime-multi-identity-size-invariant-timesformer-for-video-deepfake-detection_b200")
_spec = importlib.util.spec_from_file_location(
    "mintime_b200", os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mintime_b200"] = _mod
_spec.loader.exec_module(_mod)

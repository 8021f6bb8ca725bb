"""screed is pure python in real life too: the stand-in of oracle/shims is used as it is."""
import importlib.util
import os

_spec = importlib.util.spec_from_file_location("_orph_screed_shim", os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "shims", "screed", "__init__.py"))
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})

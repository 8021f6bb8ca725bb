"""Import the UNMODIFIED reference python (``/root/reference/orpheum``) in the build container.

Only oracle/gen_golden.py and build-container-only tests use this; it cannot run on the GPU
box (the reference checkout is not there).  Mechanism: a stub ``orpheum``
package whose ``__path__`` points at the reference sources, so that sub-modules import
without executing ``orpheum/__init__.py`` (which needs ipykernel), plus the stand-ins in
oracle/shims for khmer / sourmash / screed.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ORPHEUM_REFERENCE", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "orpheum"))


def install():
    """Make ``import orpheum.translate`` etc. resolve to the unmodified reference files."""
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)
    if "orpheum" not in sys.modules:
        pkg = types.ModuleType("orpheum")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "orpheum")]
        sys.modules["orpheum"] = pkg
    return sys.modules["orpheum"]

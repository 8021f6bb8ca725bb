"""khmer stand-in for TIMING the unmodified reference: Nodegraph is the C type of oracle/refnative (one C call per add /
get, like the pinned khmer wheel).  No module-level ``load_nodegraph``, so the reference's ``index.load_nodegraph``
(index.py:70-77) falls through to ``Nodegraph.load``.  Measurement infrastructure only."""
from _orph_refnative import Nodegraph  # noqa: F401

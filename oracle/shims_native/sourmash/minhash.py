"""sourmash.minhash.hash_murmur at native speed (oracle/refnative).  Measurement infrastructure only."""
from _orph_refnative import hash_murmur  # noqa: F401

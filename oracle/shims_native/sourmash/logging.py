"""sourmash.logging.notify stand-in (imported by the reference's compare_kmer_content.py:13)."""
import sys


def notify(s, *args, **kwargs):
    print(s.format(*args, **kwargs), file=sys.stderr)

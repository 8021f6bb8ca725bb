#!/usr/bin/env python
"""Stage the UNMODIFIED reference python so that it can be TIMED on the GPU box (bench.py --impl reference).

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Run by ``__graft_entry__.build()`` in the build container, where
``/root/reference`` exists; the outputs go to the git-ignored ``oracle/_ref/`` (which travels to the GPU box with the
snapshot, like every other built artefact) and nothing is ever copied into tracked files:

  oracle/_ref/orpheum/*.py       byte-for-byte copies of the modules on the translate / index path
                                 (``__init__.py`` is an EMPTY stub: the reference's own pulls in notebook tooling);
  oracle/_ref/_orph_refnative.so the CPython extension of oracle/refnative/refnative.c + oracle/orpheum_oracle.c: khmer's
                                 Nodegraph and sourmash's hash_murmur as native calls (the reference's two third-party
                                 wheels are absent here; the pure-python stand-ins of oracle/shims are ten times too slow
                                 to be a fair clock);
  oracle/_ref/STAGED.json        source path and sha256 of every staged file.

``activate()`` puts the staged package and the native stand-ins (oracle/shims_native) on ``sys.path``.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("ORPHEUM_REFERENCE", "/root/reference")
# the modules `orpheum translate` / `orpheum index` import (orpheum/translate.py:6-25, orpheum/index.py:1-13)
MODULES = ["translate.py", "index.py", "translate_single_seq.py", "sequence_encodings.py", "constants_translate.py",
           "constants_index.py", "create_save_summary.py", "compare_kmer_content.py", "log_utils.py"]
EXT_PATH = os.path.join(REF_DIR, "_orph_refnative.so")


def source_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "orpheum", "translate.py"))


def staged():
    return os.path.isfile(os.path.join(REF_DIR, "orpheum", "translate.py")) and os.path.isfile(EXT_PATH)


def stage(verbose=False):
    """Copies the reference modules and compiles the native stand-ins.  No-op (returns False) without /root/reference."""
    if not source_available():
        return False
    pkg = os.path.join(REF_DIR, "orpheum")
    os.makedirs(pkg, exist_ok=True)
    manifest = {"source": os.path.join(REFERENCE_ROOT, "orpheum"), "files": {}}
    for name in MODULES:
        src = os.path.join(REFERENCE_ROOT, "orpheum", name)
        shutil.copyfile(src, os.path.join(pkg, name))
        with open(src, "rb") as f:
            manifest["files"][name] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(pkg, "__init__.py"), "w") as f:
        f.write("")  # stub: the reference's __init__ imports notebook helpers that are not installed
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-pthread", "-I", sysconfig.get_paths()["include"], "-o", EXT_PATH,
           os.path.join(HERE, "refnative", "refnative.c"), os.path.join(HERE, "orpheum_oracle.c"), "-lm"]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    with open(os.path.join(REF_DIR, "STAGED.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    return True


def activate():
    """sys.path entries that make ``import orpheum.translate`` resolve to the staged reference with native khmer / sourmash."""
    if not staged():
        raise RuntimeError("oracle/_ref is not staged (run __graft_entry__.build() where /root/reference exists)")
    for p in (os.path.join(HERE, "shims_native"), REF_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)


if __name__ == "__main__":
    print("staged" if stage(verbose=True) else f"no reference checkout at {REFERENCE_ROOT}")

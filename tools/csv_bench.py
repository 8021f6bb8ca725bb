import os, sys, time, ctypes, numpy as np
sys.path.insert(0, '/root/repo')
from orpheum_b200 import _lib
lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
rng = np.random.default_rng(1)
rec = np.zeros(n, dtype=_lib.READ_RESULT_DTYPE)
rec["category"] = rng.integers(0, 6, (n, 6), dtype=np.uint8)
rec["n_kmers"] = rng.integers(1, 45, (n, 6), dtype=np.uint16)
rec["n_hits"] = (rec["n_kmers"] * rng.random((n, 6))).astype(np.uint16)
names = [b"read%d lane:1/1" % i for i in range(n)]
offs = np.zeros(n + 1, dtype=np.uint64); offs[1:] = np.cumsum([len(x) for x in names])
blob = np.frombuffer(b"".join(names), dtype=np.uint8)
texts = [b"t%d text of category" % i for i in range(6)]
arr = (ctypes.c_char_p * 6)(*texts)
path = "/dev/shm/csvbench.csv"
for rep in range(6):
    if os.path.exists(path): os.unlink(path)
    with open(path, "wb") as f: f.write(b"h\n")
    t = time.perf_counter()
    _lib.check(lib.orph_csv_write_scores(os.fsencode(path), 1, _lib.ptr(blob), _lib.ptr(offs), n, _lib.ptr(rec), arr, b"reads.fq", 0))
    dt = time.perf_counter() - t
    sz = os.path.getsize(path)
    print(f"{sz/1e9:.2f} GB in {dt:.3f} s = {sz/dt/1e9:.2f} GB/s")
import hashlib
print(hashlib.sha256(open(path,'rb').read()).hexdigest()[:16])
os.unlink(path)

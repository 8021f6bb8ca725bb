"""Seeded synthetic workloads for the parity tests and ``bench.py``.

Nothing here is on the scoring path; it only manufactures inputs of the shapes
``BASELINE.json`` names: a proteome with realistic residue frequencies, and reads that are
either back-translated from that proteome (coding) or uniform ACGT (random), plus the
low-complexity / ``N`` / short / mixed-case flavours the classification branches need.
"""
import numpy as np

AMINO_ACIDS = "SLTEAVPGKDRIQNFHYMCW"
# empirical residue frequencies of the reference's test proteome subset, in AMINO_ACIDS order
_AA_FREQ = np.array(
    [9.8, 9.4, 7.8, 6.8, 6.8, 6.4, 6.2, 6.0, 5.7, 5.2, 4.6, 4.6, 4.3, 3.5, 3.3, 2.5, 2.3, 2.2, 1.6, 0.9]
)

# standard genetic code, index = 16*b0 + 4*b1 + b2 with T,C,A,G = 0..3
_CODE_TCAG = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"
_BASES_TCAG = "TCAG"


def _synonymous_codon_table():
    """(codons[256, 6, 3] uint8, counts[256] int64): sense codons per amino-acid byte."""
    codons = np.zeros((256, 6, 3), dtype=np.uint8)
    counts = np.zeros(256, dtype=np.int64)
    for i, aa in enumerate(_CODE_TCAG):
        if aa == "*":
            continue
        trip = (_BASES_TCAG[i // 16], _BASES_TCAG[(i // 4) % 4], _BASES_TCAG[i % 4])
        a = ord(aa)
        codons[a, counts[a]] = [ord(c) for c in trip]
        counts[a] += 1
    return codons, counts


_SYN_CODONS, _SYN_COUNTS = _synonymous_codon_table()
_COMPLEMENT = np.arange(256, dtype=np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMPLEMENT[_a] = _b
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def make_proteome(n_proteins=20000, seed=1001):
    """Synthetic proteome: (residues uint8[total], offsets uint64[n+1]).  P20M: defaults; P100M:
    n_proteins=187000, seed=1004."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lengths = np.clip(np.round(rng.lognormal(6.1, 0.6, n_proteins)), 50, 5000).astype(np.int64)
    offsets = np.zeros(n_proteins + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lengths)
    p = _AA_FREQ / _AA_FREQ.sum()
    idx = rng.choice(len(AMINO_ACIDS), size=int(offsets[-1]), p=p)
    residues = np.frombuffer(AMINO_ACIDS.encode(), dtype=np.uint8)[idx]
    return np.ascontiguousarray(residues), offsets


def _coding_reads(rng, n, length, residues, prot_offsets, sub_rate):
    """n back-translated reads of `length` nt (uint8 [n, length])."""
    n_res = length // 3 + 2
    n_prot = len(prot_offsets) - 1
    plen = (prot_offsets[1:] - prot_offsets[:-1]).astype(np.int64)
    ok = np.nonzero(plen >= n_res)[0]
    prot = ok[rng.integers(0, len(ok), n)]
    start = prot_offsets[prot].astype(np.int64) + (rng.random(n) * (plen[prot] - n_res + 1)).astype(np.int64)
    aa = residues[start[:, None] + np.arange(n_res)[None, :]]          # [n, n_res]
    pick = (rng.random(aa.shape) * _SYN_COUNTS[aa]).astype(np.int64)
    nt = _SYN_CODONS[aa, pick].reshape(n, n_res * 3)                   # [n, 3*n_res]
    phase = rng.integers(0, 3, n)
    cols = phase[:, None] + np.arange(length)[None, :]
    reads = np.take_along_axis(nt, cols, axis=1)
    # substitutions: replace by one of the three other bases
    sub = rng.random(reads.shape) < sub_rate
    if sub.any():
        code = np.zeros(256, dtype=np.uint8)
        code[_ACGT] = np.arange(4, dtype=np.uint8)
        shifted = (code[reads[sub]] + rng.integers(1, 4, int(sub.sum()))) % 4
        reads[sub] = _ACGT[shifted]
    # reverse-complement half of them
    flip = rng.random(n) < 0.5
    reads[flip] = _COMPLEMENT[reads[flip][:, ::-1]]
    return reads


def make_reads_fixed(n_reads, residues, prot_offsets, length=150, seed=2002, sub_rate=0.01, chunk=1 << 18):
    """R150 recipe: read i is coding iff i is even, else uniform ACGT.  Returns uint8 [n_reads, length]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.empty((n_reads, length), dtype=np.uint8)
    for lo in range(0, n_reads, chunk):
        hi = min(n_reads, lo + chunk)
        m = hi - lo
        block = _ACGT[rng.integers(0, 4, (m, length))]
        even = np.nonzero((np.arange(lo, hi) % 2) == 0)[0]
        if len(even):
            block[even] = _coding_reads(rng, len(even), length, residues, prot_offsets, sub_rate)
        out[lo:hi] = block
    return out


def make_reads_mixed(n_reads, residues, prot_offsets, ksize=7, seed=5005, min_len=50, max_len=300):
    """cfg5 recipe: mixed lengths with every classification branch represented.
    45 % coding, 45 % random, 4 % homopolymer-rich, 3 % triplet repeats, 2 % with 1-5 N, 1 % shorter
    than 3k+3.  Returns (bytes uint8[total], offsets uint64[n+1])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    kinds = rng.choice(6, size=n_reads, p=[0.45, 0.45, 0.04, 0.03, 0.02, 0.01])
    lengths = rng.integers(min_len, max_len + 1, n_reads)
    short = kinds == 5
    lengths[short] = rng.integers(1, 3 * ksize + 3, int(short.sum()))
    offsets = np.zeros(n_reads + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lengths)
    buf = _ACGT[rng.integers(0, 4, int(offsets[-1]))]
    coding_full = _coding_reads(rng, int((kinds == 0).sum()) + int((kinds == 4).sum()), max_len, residues,
                                prot_offsets, 0.01)
    ci = 0
    for i in range(n_reads):
        lo, L, kind = int(offsets[i]), int(lengths[i]), int(kinds[i])
        if kind == 0 or kind == 4:
            buf[lo:lo + L] = coding_full[ci, :L]
            ci += 1
            if kind == 4:
                pos = rng.integers(0, L, rng.integers(1, 6))
                buf[lo + pos] = ord("N")
        elif kind == 2:
            # long runs of one base with sparse interruptions: ndiff / L < 0.3
            base = _ACGT[rng.integers(0, 4)]
            seg = np.full(L, base, dtype=np.uint8)
            n_breaks = int(0.1 * L)
            if n_breaks:
                seg[rng.integers(0, L, n_breaks)] = _ACGT[rng.integers(0, 4, n_breaks)]
            buf[lo:lo + L] = seg
        elif kind == 3:
            unit = _ACGT[rng.integers(0, 4, 3)]
            buf[lo:lo + L] = np.tile(unit, L // 3 + 1)[:L]
    return buf, offsets


def concat_fixed(reads2d):
    """[n, L] matrix -> (flat bytes, offsets) in the layout the C ABI takes."""
    n, L = reads2d.shape
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
    return np.ascontiguousarray(reads2d).reshape(-1), offsets


# ---- counter-based reads (cfg4): read i is a pure function of (seed, i) --------------------------------------
# numpy twin of csrc/synth_kernels.cuh; both must produce the same bases for every index.
_GOLD, _STEP = np.uint64(0x9E3779B97F4A7C15), np.uint64(0xD1B54A32D192ED03)
_CODE_OF = np.zeros(256, dtype=np.uint8)          # ASCII -> 2-bit read code (ascii >> 1) & 3: A0 C1 T2 G3
_CODE_OF[_ACGT] = (_ACGT >> 1) & 3
_BASE_OF = np.frombuffer(b"ACTG", dtype=np.uint8)  # 2-bit read code -> ASCII


def syn_table():
    """uint64[256] synonymous-codon table in the layout orph_synth_reads_device takes: bits 0..2 = number of sense
    codons of the amino-acid byte, codon c at bits [4 + 6c, +6) as three 2-bit read codes (first base lowest)."""
    table = np.zeros(256, dtype=np.uint64)
    for aa in range(256):
        n = int(_SYN_COUNTS[aa])
        v = n
        for c in range(n):
            codes = _CODE_OF[_SYN_CODONS[aa, c]]
            v |= (int(codes[0]) | (int(codes[1]) << 2) | (int(codes[2]) << 4)) << (4 + 6 * c)
        table[aa] = v
    return table


def _mix(z):
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _draw(seed, idx, d):
    with np.errstate(over="ignore"):
        return _mix(np.uint64(seed) + idx * _GOLD + np.uint64(d + 1) * _STEP)


def counter_reads(indices, residues, prot_offsets, seed=3003, length=150):
    """Reads with the given GLOBAL indices: uint8 [m, length] ASCII.  Even index: back-translated; odd: uniform."""
    idx = np.ascontiguousarray(indices, dtype=np.uint64)
    m, n_res = len(idx), length // 3 + 2
    codes = np.zeros((m, length), dtype=np.uint8)
    with np.errstate(over="ignore"):
        odd = (idx & np.uint64(1)) == 1
        b = np.arange(length)
        if odd.any():
            draws = np.stack([_draw(seed, idx[odd], q) for q in range((length + 31) // 32)], axis=1)  # [mo, words]
            codes[odd] = ((draws[:, b // 32] >> (2 * (b % 32)).astype(np.uint64)) & np.uint64(3)).astype(np.uint8)
        ev = ~odd
        if ev.any():
            ie = idx[ev]
            n_prot = len(prot_offsets) - 1
            plen = (prot_offsets[1:] - prot_offsets[:-1]).astype(np.uint64)
            d0, d1 = _draw(seed, ie, 0), _draw(seed, ie, 1)
            p = (d0 % np.uint64(n_prot)).astype(np.int64)
            while True:
                short = plen[p] < n_res
                if not short.any():
                    break
                p[short] = (p[short] + 1) % n_prot
            start = prot_offsets[p].astype(np.uint64) + (d0 >> np.uint64(32)) % (plen[p] - np.uint64(n_res) + np.uint64(1))
            phase = (d1 % np.uint64(3)).astype(np.int64)
            flip = ((d1 >> np.uint64(8)) & np.uint64(1)).astype(bool)
            j = np.arange(n_res)
            aa = residues[start.astype(np.int64)[:, None] + j[None, :]]                               # [me, n_res]
            picks = np.stack([_draw(seed, ie, 2 + q) for q in range((n_res + 7) // 8)], axis=1)       # [me, groups]
            pick = ((picks[:, j // 8] >> (8 * (j % 8)).astype(np.uint64)) & np.uint64(0xFF)).astype(np.int64)
            choice = pick % _SYN_COUNTS[aa]
            nt = _CODE_OF[_SYN_CODONS[aa, choice].reshape(len(ie), n_res * 3)]                        # 2-bit codes
            rd = np.take_along_axis(nt, phase[:, None] + b[None, :], axis=1)
            subs = np.stack([_draw(seed, ie, 16 + q) for q in range((length + 3) // 4)], axis=1)
            v = ((subs[:, b // 4] >> (16 * (b % 4)).astype(np.uint64)) & np.uint64(0xFFFF)).astype(np.int64)
            hit = v < 655
            rd = np.where(hit, (rd + 1 + v % 3) & 3, rd).astype(np.uint8)
            rd[flip] = rd[flip][:, ::-1] ^ 2
            codes[ev] = rd
    return _BASE_OF[codes]


def write_bgzf(path, data, block=65280, level=1, eof_marker=True):
    """Blocked gzip (BGZF, SAM spec 4.1): independent members of at most 64 KiB with a "BC" extra field that holds the
    member's size -- what bgzip / samtools and many sequencing pipelines write."""
    import struct
    import zlib

    with open(path, "wb") as f:
        chunks = [data[i:i + block] for i in range(0, len(data), block)] + ([b""] if eof_marker else [])
        for chunk in chunks:
            comp = zlib.compressobj(level, zlib.DEFLATED, -15)
            body = comp.compress(chunk) + comp.flush()
            f.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, len(body) + 25))
            f.write(body + struct.pack("<II", zlib.crc32(chunk), len(chunk)))

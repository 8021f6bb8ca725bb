"""khmer.Nodegraph stand-in (pinned khmer==3.0.0a3 / 2.1.1 in the reference's requirement files).

Deliberately has NO module-level ``load_nodegraph`` so that the reference's
``index.load_nodegraph`` (index.py:70-77) falls through to ``Nodegraph.load``.
"""
import struct


def _is_prime(n):
    if n < 2:
        return False
    if n < 4:
        return True
    if n % 2 == 0:
        return False
    d = 3
    while d * d <= n:
        if n % d == 0:
            return False
        d += 2
    return True


def get_n_primes_near_x(n, x):
    primes = []
    i = int(x) - 1
    if i % 2 == 0:
        i -= 1
    while len(primes) != n and i > 0:
        if _is_prime(i):
            primes.append(i)
        i -= 2
    return primes


class Nodegraph:
    def __init__(self, k, starting_size, n_tables, primes=None):
        self._k = int(k)
        self._sizes = list(primes) if primes else get_n_primes_near_x(int(n_tables), int(starting_size))
        self._tables = [bytearray(s // 8 + 1) for s in self._sizes]
        self._n_unique = 0

    def ksize(self):
        return self._k

    def hashsizes(self):
        return list(self._sizes)

    def n_tables(self):
        return len(self._sizes)

    def n_unique_kmers(self):
        return self._n_unique

    def n_occupied(self):
        return sum(bin(b).count("1") for b in self._tables[0])

    def add(self, h):
        is_new = False
        for size, table in zip(self._sizes, self._tables):
            b = h % size
            mask = 1 << (b & 7)
            if not table[b >> 3] & mask:
                table[b >> 3] |= mask
                is_new = True
        if is_new:
            self._n_unique += 1
        return is_new

    def get(self, h):
        for size, table in zip(self._sizes, self._tables):
            b = h % size
            if not table[b >> 3] & (1 << (b & 7)):
                return 0
        return 1

    def save(self, path):
        with open(path, "wb") as f:
            f.write(b"OXLI" + bytes([4, 2]) + struct.pack("<IBQ", self._k, len(self._sizes), self.n_occupied()))
            for size, table in zip(self._sizes, self._tables):
                f.write(struct.pack("<Q", size))
                f.write(bytes(table))

    @classmethod
    def load(cls, path):
        with open(path, "rb") as f:
            head = f.read(19)
            if len(head) != 19 or head[:4] != b"OXLI" or head[4] != 4 or head[5] != 2:
                raise OSError(f"{path} is not an OXLI nodegraph file")
            k, n_tables, _occupied = struct.unpack("<IBQ", head[6:])
            sizes, tables = [], []
            for _ in range(n_tables):
                (size,) = struct.unpack("<Q", f.read(8))
                data = f.read(size // 8 + 1)
                if len(data) != size // 8 + 1:
                    raise OSError(f"{path} is truncated")
                sizes.append(size)
                tables.append(bytearray(data))
        g = cls(k, 0, n_tables, primes=sizes)
        g._tables = tables
        return g

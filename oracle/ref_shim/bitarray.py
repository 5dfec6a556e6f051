"""Minimal pure-Python stand-in for the `bitarray` package (not installed, no
network).  TEST INFRASTRUCTURE: lets oracle/ref_harness.py import the UNMODIFIED
reference (adaptive_compressor.py:10, marker_finder.py:3).  Implements only what
those files use: bitarray(), bitarray(str), frombytes, to01, tobytes, append,
len, slicing, + (big-endian / MSB-first bit order, the package default)."""


class bitarray:
    def __init__(self, init=None):
        self._b = []
        if isinstance(init, str):
            self._b = [1 if c == "1" else 0 for c in init]
        elif isinstance(init, bitarray):
            self._b = list(init._b)
        elif init is not None:
            self._b = [1 if x else 0 for x in init]

    def frombytes(self, data):
        for byte in data:
            for k in range(7, -1, -1):
                self._b.append((byte >> k) & 1)

    def to01(self):
        return "".join("1" if x else "0" for x in self._b)

    def tobytes(self):
        bits = self._b + [0] * ((-len(self._b)) % 8)
        out = bytearray()
        for i in range(0, len(bits), 8):
            v = 0
            for x in bits[i:i + 8]:
                v = (v << 1) | x
            out.append(v)
        return bytes(out)

    def append(self, x):
        self._b.append(1 if x else 0)

    def __len__(self):
        return len(self._b)

    def __getitem__(self, i):
        if isinstance(i, slice):
            r = bitarray()
            r._b = self._b[i]
            return r
        return self._b[i]

    def __add__(self, other):
        r = bitarray()
        r._b = self._b + other._b
        return r

"""Restatement of the seeded RNG behind the reference's snapshot tests — test infrastructure only.

The reference's legacy tests (spenso/src/tests.rs:57-148, 502-532) build their inputs with
`rand_xoshiro::Xoroshiro64Star::seed_from_u64(seed)` and `rand` distributions.  Those crates are
third-party and un-vendored (Cargo.lock pins rand 0.8.5, rand_core 0.6.4, rand_xoshiro 0.6.0), so
their PUBLISHED algorithms are restated here:

* SplitMix64 (Steele/Lea/Flood) expands the u64 seed into the 8 seed bytes;
* xoroshiro64* (Blackman/Vigna): output s0 * 0x9E3779BB, state update rotl 26 / shl 9 / rotl 13;
  a u64 is two u32 draws, low word first;
* rand 0.8.5 `UniformInt`: widening multiply + rejection zone (`sample_single_inclusive` for
  `gen_range`, the precomputed-zone variant for `Uniform::new(..).sample`); i8/i16 sample through u32;
* `Standard` for i8: low byte of one u32 draw.

Pinned: with these, `rng_is_deterministic`'s snapshot (15 numbers), the structures of five other
snapshots, the 79 of `trace`, the 40 numbers of `arithmetic_data` and the contraction results of
`scalar_and_dim1_conract` / `multi_contract` are reproduced exactly (tests/test_reference_snapshots.py).
"""
from __future__ import annotations

M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF


def _rotl32(x, k):
    return ((x << k) | (x >> (32 - k))) & M32


class SplitMix64:
    def __init__(self, seed: int):
        self.x = seed & M64

    def next_u64(self) -> int:
        self.x = (self.x + 0x9E3779B97F4A7C15) & M64
        z = self.x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
        return z ^ (z >> 31)


class Xoroshiro64Star:
    def __init__(self, seed: int):
        v = SplitMix64(seed).next_u64()          # seed_from_u64: 8 little-endian seed bytes
        self.s0, self.s1 = v & M32, v >> 32
        assert self.s0 or self.s1

    def next_u32(self) -> int:
        r = (self.s0 * 0x9E3779BB) & M32
        self.s1 ^= self.s0
        self.s0 = _rotl32(self.s0, 26) ^ self.s1 ^ ((self.s1 << 9) & M32)
        self.s1 = _rotl32(self.s1, 13)
        return r

    def next_u64(self) -> int:
        lo = self.next_u32()
        hi = self.next_u32()
        return (hi << 32) | lo

    # ---- rand 0.8.5 UniformInt::sample_single_inclusive (gen_range) ----------------------------
    def _single_inclusive(self, low: int, high: int, bits: int) -> int:
        """bits = width of the unsigned sampling type (32 for i32, 64 for usize)."""
        mask = (1 << bits) - 1
        span = (high - low + 1) & mask
        assert span != 0
        zone = (((span << (bits - span.bit_length())) & mask) - 1) & mask
        while True:
            v = self.next_u32() if bits == 32 else self.next_u64()
            m = v * span
            if (m & mask) <= zone:
                return low + (m >> bits)

    def gen_range_i32(self, low: int, high: int) -> int:            # low..high
        return self._single_inclusive(low, high - 1, 32)

    def gen_range_incl_i32(self, low: int, high: int) -> int:       # low..=high
        return self._single_inclusive(low, high, 32)

    def gen_range_usize(self, low: int, high: int) -> int:
        return self._single_inclusive(low, high - 1, 64)

    def gen_range_incl_usize(self, low: int, high: int) -> int:
        return self._single_inclusive(low, high, 64)

    # ---- Standard ---------------------------------------------------------------------------------
    def gen_i8(self) -> int:
        v = self.next_u32() & 0xFF
        return v - 256 if v >= 128 else v

    # ---- Uniform::new(low, high).sample for i16 / i32 (both sample through u32) ---------------------
    def uniform(self, low: int, high: int, ty_bits: int) -> int:
        span = (high - 1 - low + 1) & ((1 << ty_bits) - 1)
        reject = ((M32 - span + 1) % span) & ((1 << ty_bits) - 1)
        zone = M32 - reject
        while True:
            m = self.next_u32() * span
            if (m & M32) <= zone:
                return low + (m >> 32)


# ---- the reference's test helpers (tests.rs), as slot tuples (rep name, dim, aind) --------------------
def test_structure(length: int, seed: int):
    """tests.rs:92-110: `length` distinct random euc/mink slots (IndexSet: insertion order)."""
    r, s = Xoroshiro64Star(seed), []
    while len(s) < length:
        rep = r.gen_range_incl_i32(0, 1)
        dim = r.gen_range_incl_usize(1, 9)
        aind = r.gen_range_i32(0, 256)
        sl = ("euc" if rep == 0 else "mink", dim, aind)
        if sl not in s:
            s.append(sl)
    return s


def test_structure_with_dims(dims, seed: int):
    """tests.rs:112-147: d < 0 -> euc | lor of dim -d ; d >= 0 -> euc | dual lor of dim d."""
    r, s = Xoroshiro64Star(seed), []
    for d in dims:
        while True:
            rep = r.gen_range_incl_i32(0, 1)
            aind = r.gen_range_i32(0, 256)
            sl = ("euc" if rep == 0 else ("lor" if d < 0 else "lor_d"), abs(d), aind)
            if sl not in s:
                s.append(sl)
                break
    return s


def test_structure_with_id(ids, seed: int):
    """tests.rs:502-532: id > 0 -> euc | lor ; id <= 0 -> euc | dual lor, index -id."""
    r, s = Xoroshiro64Star(seed), []
    for i in ids:
        rep = r.gen_range_incl_i32(0, 1)
        dim = r.gen_range_incl_usize(1, 9)
        s.append(("euc" if rep == 0 else ("lor" if i > 0 else "lor_d"), dim, abs(i)))
    return s


def test_tensor(size: int, seed: int, kind: str, low=None, high=None) -> dict:
    """tests.rs:57-90: {flat index: value}; later writes overwrite earlier ones (HashMap insert).
    kind: 'i8' (Standard, range None) | 'i16' | 'i32' (Uniform::new(low, high))."""
    r, d = Xoroshiro64Star(seed), {}
    density = r.gen_range_usize(0, size)
    for _ in range(density):
        i = r.gen_range_usize(0, size)
        d[i] = r.gen_i8() if kind == "i8" else r.uniform(low, high, 16 if kind == "i16" else 32)
    return d

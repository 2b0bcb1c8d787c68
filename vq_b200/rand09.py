"""Host-side index stream for PQ training: a restatement of the `rand 0.9` calls the
reference makes at src/core/vector.rs:412-413 and :450,

    let mut rng = StdRng::seed_from_u64(seed);
    data.choose_multiple(&mut rng, k)          // initial centroids
    data.choose(&mut rng)                      // re-seed of an empty cluster

PARITY UNPINNED.  The `rand`, `rand_chacha` and `rand_core` crates are third-party
dependencies that are NOT vendored in the reference checkout (Cargo.toml:41, Cargo.lock
is git-ignored), no Rust toolchain exists in this environment, and no reference test
pins a sampled index.  This module restates the published 0.9 algorithms (ChaCha12
StdRng, PCG32 seed expansion, index::sample with its Floyd / in-place / rejection
branches, Canon's method for single-sample integer ranges) as faithfully as they are
known, but it cannot be verified here; the engine therefore takes the index stream as
an explicit input (init_idx + reseed callback, include/vqb200.h) and parity with the
reference is asserted conditional on that stream.  Both calls depend only on the slice
LENGTH, so a verified Rust shim can feed the same entry points from the real crate.
"""
from __future__ import annotations

import numpy as np

_M32 = 0xFFFFFFFF
_M64 = 0xFFFFFFFFFFFFFFFF


def _rotl32(v, c):
    return ((v << c) & _M32) | (v >> (32 - c))


def _chacha_block(key_words, counter, stream, rounds=12):
    c = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574]
    st = c + list(key_words) + [counter & _M32, (counter >> 32) & _M32, stream & _M32, (stream >> 32) & _M32]
    x = list(st)

    def qr(a, b, cc, d):
        x[a] = (x[a] + x[b]) & _M32; x[d] = _rotl32(x[d] ^ x[a], 16)
        x[cc] = (x[cc] + x[d]) & _M32; x[b] = _rotl32(x[b] ^ x[cc], 12)
        x[a] = (x[a] + x[b]) & _M32; x[d] = _rotl32(x[d] ^ x[a], 8)
        x[cc] = (x[cc] + x[d]) & _M32; x[b] = _rotl32(x[b] ^ x[cc], 7)

    for _ in range(rounds // 2):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(x[i] + st[i]) & _M32 for i in range(16)]


class StdRng:
    """rand 0.9 StdRng: ChaCha12, 64-bit block counter, stream 0, 4-block output buffer."""

    def __init__(self, key_words):
        self.key = list(key_words)
        self.counter = 0
        self.buf = []
        self.index = 64  # empty

    @classmethod
    def seed_from_u64(cls, state: int) -> "StdRng":
        # rand_core SeedableRng::seed_from_u64: PCG32 output, 8 little-endian words
        state &= _M64
        words = []
        for _ in range(8):
            state = (state * 6364136223846793005 + 11634580027462260723) & _M64
            xorshifted = (((state >> 18) ^ state) >> 27) & _M32
            rot = state >> 59
            words.append(((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & _M32)
        return cls(words)

    def _generate(self, index):
        self.buf = []
        for b in range(4):
            self.buf += _chacha_block(self.key, self.counter + b, 0)
        self.counter += 4
        self.index = index

    def next_u32(self) -> int:
        if self.index >= 64:
            self._generate(0)
        v = self.buf[self.index]
        self.index += 1
        return v

    def next_u64(self) -> int:  # rand_core BlockRng::next_u64
        if self.index < 63:
            if not self.buf:
                self._generate(0)
            lo, hi = self.buf[self.index], self.buf[self.index + 1]
            self.index += 2
            return (hi << 32) | lo
        if self.index >= 64:
            self._generate(2)
            return (self.buf[1] << 32) | self.buf[0]
        lo = self.buf[63]
        self._generate(1)
        return (self.buf[0] << 32) | lo

    # ---- uniform integers -------------------------------------------------
    def _random_below_single(self, rng_range: int, wide: bool) -> int:
        """UniformInt::sample_single_inclusive for [0, range-1] (Canon's method, one extra draw)."""
        bits = 64 if wide else 32
        mask = _M64 if wide else _M32
        draw = self.next_u64 if wide else self.next_u32
        if rng_range == 0:  # full range
            return draw()
        prod = draw() * rng_range
        result, lo = prod >> bits, prod & mask
        if lo > ((-rng_range) & mask):
            new_hi = (draw() * rng_range) >> bits
            if lo + new_hi > mask:
                result += 1
        return result

    def random_range_usize(self, n: int) -> int:
        """rng.random_range(..n) for usize: sampled as u32 when n fits (0.9 portability rule)."""
        if n <= 0:
            raise ValueError("empty range")
        return self._random_below_single(n, wide=n > _M32)

    def _uniform_u32(self, length: int) -> int:
        """Uniform::<u32>::new(0, length).sample(rng): widening multiply with exact rejection zone."""
        thresh = ((-length) & _M32) % length
        while True:
            prod = self.next_u32() * length
            if (prod & _M32) >= thresh:
                return prod >> 32

    # ---- rand::seq::index::sample -----------------------------------------
    def sample_indices(self, length: int, amount: int):
        if amount > length:
            raise ValueError("`amount` of samples must be less than or equal to `length`")
        if length > _M32:
            return self._sample_rejection(length, amount, wide=True)
        if amount < 163:
            c = [[1.6, 8.0 / 45.0], [10.0, 70.0 / 9.0]]
            j = 1 if length >= 500_000 else 0
            amount_fp = np.float32(amount)
            m4 = np.float32(c[0][j]) * amount_fp
            if amount > 11 and np.float32(length) < (np.float32(c[1][j]) + m4) * amount_fp:
                return self._sample_inplace(length, amount)
            return self._sample_floyd(length, amount)
        c = [270.0, 330.0 / 9.0]
        j = 1 if length >= 500_000 else 0
        if np.float32(length) < np.float32(c[j]) * np.float32(amount):
            return self._sample_inplace(length, amount)
        return self._sample_rejection(length, amount, wide=False)

    def _sample_floyd(self, length, amount):
        indices = []
        for j in range(length - amount, length):
            t = self._random_below_single((j + 1) & _M32, wide=False)  # random_range(..=j)
            if t in indices:
                indices[indices.index(t)] = j
            indices.append(t)
        return indices

    def _sample_inplace(self, length, amount):
        indices = list(range(length))
        for i in range(amount):
            j = i + self._random_below_single(length - i, wide=False)  # random_range(i..length)
            indices[i], indices[j] = indices[j], indices[i]
        return indices[:amount]

    def _sample_rejection(self, length, amount, wide):
        seen, out = set(), []
        for _ in range(amount):
            while True:
                pos = self._uniform_u32(length) if not wide else self._uniform_u64(length)
                if pos not in seen:
                    seen.add(pos)
                    out.append(pos)
                    break
        return out

    def _uniform_u64(self, length):
        thresh = ((-length) & _M64) % length
        while True:
            prod = self.next_u64() * length
            if (prod & _M64) >= thresh:
                return prod >> 64


class IndexStream:
    """Per-subspace stream: StdRng::seed_from_u64(seed + subspace) (src/pq.rs:130)."""

    def __init__(self, seed: int, subspace: int):
        self.rng = StdRng.seed_from_u64((seed + subspace) & _M64)

    def choose_multiple(self, n: int, k: int):
        """data.choose_multiple(&mut rng, k): k distinct indices in algorithm order."""
        return self.rng.sample_indices(n, min(k, n))

    def choose(self, n: int) -> int:
        """data.choose(&mut rng)."""
        return self.rng.random_range_usize(n)

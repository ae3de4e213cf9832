"""Oracle (test infrastructure): tube masking with the MOFO motion-box constraint.

Restates ``/root/reference/masking_generator.py``:
  * ``TubeMaskingGenerator.__call__``      masking_generator.py:17-24
  * ``TubeMaskingGenerator_BB.__call__``   masking_generator.py:43-85

The reference draws from numpy's *legacy global* RandomState (MT19937).  To make
"same random draw" a testable contract, every function here consumes an explicit
array of raw 32-bit MT19937 outputs (``rng_words``) — exactly the words
``np.random.RandomState(seed)._bit_generator.random_raw(n)`` would return, which are the
words ``np.random.shuffle`` consumes after ``np.random.seed(seed)``.  A self-contained
MT19937 (``mt19937_words``) is included so the oracle does not depend on numpy
internals; ``tests/test_oracle_mask.py`` pins it against numpy and against the reference.

numpy legacy shuffle (numpy/random/mtrand.pyx ``shuffle`` -> ``random_interval``
in numpy/random/src/legacy/legacy-distributions.c / distributions.c, numpy 1.x-2.x,
stream frozen by NEP 19):
    for i in n-1 .. 1:  j = interval(i); swap(x[i], x[j])
    interval(max): mask = smallest (2^k - 1) >= max; draw 32-bit words w until
                   (w & mask) <= max; return it.  (max == 0 draws nothing.)
"""
from __future__ import annotations

import numpy as np

PATCH = 16  # hard-coded in the reference, masking_generator.py:50-53


def mt19937_words(seed: int, n: int) -> np.ndarray:
    """First ``n`` raw 32-bit outputs of MT19937 seeded like ``np.random.seed(seed)``
    (init_genrand; numpy/random/_mt19937.pyx ``_legacy_seeding`` for an int seed)."""
    mt = np.zeros(624, dtype=np.uint64)
    mt[0] = seed & 0xFFFFFFFF
    for i in range(1, 624):
        mt[i] = (1812433253 * (int(mt[i - 1]) ^ (int(mt[i - 1]) >> 30)) + i) & 0xFFFFFFFF
    state = [int(v) for v in mt]
    out = np.empty(n, dtype=np.uint32)
    idx = 624
    for k in range(n):
        if idx >= 624:
            for i in range(624):
                y = (state[i] & 0x80000000) | (state[(i + 1) % 624] & 0x7FFFFFFF)
                v = state[(i + 397) % 624] ^ (y >> 1)
                if y & 1:
                    v ^= 0x9908B0DF
                state[i] = v
            idx = 0
        y = state[idx]
        idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        out[k] = y & 0xFFFFFFFF
    return out


class _WordStream:
    def __init__(self, words):
        self.words = np.asarray(words, dtype=np.uint32)
        self.pos = 0

    def interval(self, mx: int) -> int:
        if mx == 0:
            return 0
        mask = mx
        mask |= mask >> 1
        mask |= mask >> 2
        mask |= mask >> 4
        mask |= mask >> 8
        mask |= mask >> 16
        while True:
            if self.pos >= len(self.words):
                raise ValueError("rng_words exhausted")
            v = int(self.words[self.pos]) & mask
            self.pos += 1
            if v <= mx:
                return v


def legacy_shuffle(x: list, ws: _WordStream) -> None:
    """In-place numpy-legacy Fisher-Yates (see module docstring)."""
    for i in range(len(x) - 1, 0, -1):
        j = ws.interval(i)
        x[i], x[j] = x[j], x[i]


def in_box_indices(bb0, height: int, width: int) -> list:
    """masking_generator.py:46-57.  ``bb0`` = first frame's box (x1,y1,x2,y2) in pixels.
    Quirks preserved: the predicate is a *cross* (x-overlap
    OR y-overlap) and box x is compared against the tube ROW index j."""
    b0, b1, b2, b3 = (float(v) for v in bb0)
    idx = []
    for j in range(height):
        for k in range(width):
            x1t, x2t = j * PATCH, j * PATCH + PATCH
            y1t, y2t = k * PATCH, k * PATCH + PATCH
            if not ((b0 > x2t or b2 < x1t) and (b1 > y2t or b3 < y1t)):
                idx.append(j * width + k)
    return idx


def tube_mask_bb(bb, rng_words, input_size=(8, 14, 14), mask_ratio=0.9, mask_ratio_bb=0.75):
    """Returns (mask float64[T*H*W] of {0.,1.}, words_used).  masking_generator.py:43-85."""
    frames, height, width = input_size
    npf = height * width
    nmask = int(mask_ratio * npf)                              # :32
    bb = np.asarray(bb, dtype=np.float64).reshape(-1, 4)
    index = in_box_indices(bb[0], height, width)               # :46-57 (frame 0 only)
    ws = _WordStream(rng_words)
    legacy_shuffle(index, ws)                                  # :62
    cap = min(nmask, int(len(index) * mask_ratio_bb))          # :64
    selected = index[:cap]
    f = np.zeros(npf)
    for i in selected:                                         # :67-68
        f[i] = 1
    remaining_masks = nmask - len(selected)                    # :71
    sel = set(selected)
    remaining = [i for i in range(nmask) if i not in sel]      # :72 setdiff1d(arange(nmask), selected) (sorted)
    legacy_shuffle(remaining, ws)                              # :75
    for i in remaining[:remaining_masks]:                      # :76-77
        f[i] = 1
    return np.tile(f, (frames, 1)).flatten(), ws.pos           # :84


def tube_mask_plain(rng_words, input_size=(8, 14, 14), mask_ratio=0.9):
    """masking_generator.py:17-24 (non-BB generator)."""
    frames, height, width = input_size
    npf = height * width
    nmask = int(mask_ratio * npf)
    m = [0.0] * (npf - nmask) + [1.0] * nmask
    ws = _WordStream(rng_words)
    legacy_shuffle(m, ws)
    return np.tile(np.asarray(m), (frames, 1)).flatten(), ws.pos


def index_lists(mask_row: np.ndarray):
    """Ascending visible / masked token ids of one clip (what ``x[~mask]`` / ``x[mask]``
    enumerate, modeling_pretrain.py:90,261-262)."""
    m = np.asarray(mask_row).astype(bool)
    return np.nonzero(~m)[0].astype(np.int32), np.nonzero(m)[0].astype(np.int32)

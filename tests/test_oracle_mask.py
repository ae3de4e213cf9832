"""Pins oracle/mask_oracle.py against the reference's own outputs (tests/golden/mask_golden.json,
made by tests/golden/make_golden.py from /root/reference/masking_generator.py) and against numpy."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import mask_oracle as mo


@pytest.fixture(scope="module")
def gold(golden_dir):
    with open(os.path.join(golden_dir, "mask_golden.json")) as f:
        return json.load(f)


def test_mt19937_matches_golden_and_numpy(gold):
    assert mo.mt19937_words(10, 8).tolist() == gold["mt19937_seed10_first8"]
    for seed in (0, 1, 10, 4321, 2**31 + 5):
        ref = np.random.RandomState(seed)._bit_generator.random_raw(1300)
        assert np.array_equal(mo.mt19937_words(seed, 1300), ref.astype(np.uint32))


def test_legacy_shuffle_matches_numpy():
    for seed in range(20):
        n = 3 + seed * 11
        words = mo.mt19937_words(seed, 4 * n + 64)
        x = list(range(n))
        mo.legacy_shuffle(x, mo._WordStream(words))
        np.random.seed(seed)
        y = list(range(n))
        np.random.shuffle(y)
        assert x == y


def test_bb_golden_cases(gold):
    for c in gold["bb_cases"]:
        words = mo.mt19937_words(c["seed"], 700)
        bb = np.tile(np.asarray([c["box"]], dtype=np.float64), (16, 1))
        m, used = mo.tube_mask_bb(bb, words)
        u8 = m.astype(np.uint8)
        assert hashlib.sha1(u8.tobytes()).hexdigest() == c["sha1"], c
        assert np.nonzero(u8[:196] == 0)[0].tolist() == c["vis_slab0"]
        assert int(u8.sum()) == c["n_masked"] == 1408
        assert 0 < used <= 700


def test_plain_golden(gold):
    m, _ = mo.tube_mask_plain(mo.mt19937_words(gold["plain"]["seed"], 700))
    u8 = m.astype(np.uint8)
    assert hashlib.sha1(u8.tobytes()).hexdigest() == gold["plain"]["sha1"]
    assert np.nonzero(u8[:196] == 0)[0].tolist() == gold["plain"]["vis_slab0"]


def test_tiny_grid_cases(gold):
    for c in gold["tiny_grid_cases"]:
        bb = np.tile(np.asarray([c["box"]], dtype=np.float64), (16, 1))
        m, _ = mo.tube_mask_bb(bb, mo.mt19937_words(c["seed"], 200), (8, 4, 4))
        assert m.astype(np.uint8).tolist() == c["mask"]


def test_properties_random_boxes():
    rng = np.random.default_rng(0)
    for it in range(200):
        x1, y1 = rng.uniform(0, 200, 2)
        w, h = rng.uniform(1, 224, 2)
        bb = np.tile(np.asarray([[x1, y1, min(224, x1 + w), min(224, y1 + h)]]), (16, 1))
        m, used = mo.tube_mask_bb(bb, mo.mt19937_words(it, 800))
        m = m.reshape(8, 196)
        assert (m.sum(1) == 176).all()          # exactly 176 per slab
        assert (m == m[0]).all()                # tube mask: identical across slabs
        vis, msk = mo.index_lists(m.reshape(-1))
        assert len(vis) == 160 and len(msk) == 1408
        assert (np.diff(vis) > 0).all() and (np.diff(msk) > 0).all()

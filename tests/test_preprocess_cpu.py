"""SURVEY.md 8(f4), host part: BucketBatchSampler policy against batches produced by the reference's sampler."""
import os

import numpy as np

from texocr_b200.preprocess import BucketBatcher, bucket_batches

HERE = os.path.dirname(os.path.abspath(__file__))


def _rows(flat):
    return [[int(v) for v in row if v >= 0] for row in flat]


def test_bucket_batches_match_reference_sampler():
    g = np.load(os.path.join(HERE, "golden", "golden_prep_v1.npz"))
    sizes = [tuple(s) for s in g["bucket_sizes"]]
    assert bucket_batches(sizes, 4, keep_small=True) == _rows(g["bucket_plain_e0"]) == _rows(g["bucket_plain_e1"])
    assert bucket_batches(sizes, 4, keep_small=False) == _rows(g["bucket_drop_e0"])
    b = BucketBatcher(sizes, 4, keep_small=True, shuffle=True, seed=3)
    assert list(b) == _rows(g["bucket_shuf_e0"])
    assert list(b) == _rows(g["bucket_shuf_e1"])          # the seed advanced: another order
    assert _rows(g["bucket_shuf_e0"]) != _rows(g["bucket_shuf_e1"])
    assert len(b) == int(g["bucket_shuf_len"]) and len(BucketBatcher(sizes, 4, keep_small=False)) == int(g["bucket_drop_len"])
    # every batch holds one image size only; every image appears exactly once when small batches are kept
    for batch in bucket_batches(sizes, 4, keep_small=True):
        assert len({sizes[i] for i in batch}) == 1
    assert sorted(i for batch in bucket_batches(sizes, 4, keep_small=True) for i in batch) == list(range(len(sizes)))
    # defaults are the reference sampler's (data_wrangling/dataset.py:282): keep_small=False, seed=42
    assert bucket_batches(sizes, 4) == _rows(g["bucket_drop_e0"])
    assert list(BucketBatcher(sizes, 4, shuffle=True)) == bucket_batches(sizes, 4, shuffle=True, seed=42)
    assert bucket_batches([], 4) == []

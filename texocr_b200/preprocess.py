"""Input side of the path: the reference's batching policy and image transform.

* ``bucket_batches`` / ``BucketBatcher`` -- ``BucketBatchSampler`` (data_wrangling/dataset.py:280-327): images are grouped by
  identical (width, height), every group is cut into batches of ``batch_size``, optionally shuffled with Python's
  ``random`` seeded per epoch.  Pure host logic.
* ``to_model_input`` -- the deterministic part of ``img_transform`` (data_wrangling/dataset.py:365-371) on the GPU through
  ``texocr_preprocess_u8``: uint8 images in, float32 (1, H, W) device tensors out, ready for ``model.generate``.
"""
import random
from collections import OrderedDict
from typing import Iterator, List, Sequence, Tuple


def bucket_batches(sizes: Sequence[Tuple[int, int]], batch_size: int, keep_small: bool = False, shuffle: bool = False,
                   seed: int = 42) -> List[List[int]]:
    """One epoch of BucketBatchSampler.__iter__ (data_wrangling/dataset.py:303-318) over ``sizes[i]`` = (w, h) of image i."""
    groups: "OrderedDict[Tuple[int, int], List[int]]" = OrderedDict()
    for i, s in enumerate(sizes):
        groups.setdefault((int(s[0]), int(s[1])), []).append(i)            # dataset order, like ImageDataset.sizes
    batches = []
    for ids in groups.values():
        for i in range(0, len(ids), batch_size):
            batch = ids[i:i + batch_size]
            if len(batch) == batch_size or keep_small:
                batches.append(batch)
    if shuffle:
        random.seed(seed)
        random.shuffle(batches)
    return batches


class BucketBatcher:
    """Stateful form: every ``iter()`` is one epoch; a shuffling batcher advances its seed per epoch like the reference."""

    def __init__(self, sizes: Sequence[Tuple[int, int]], batch_size: int, keep_small: bool = False, shuffle: bool = False, seed: int = 42):
        # defaults as BucketBatchSampler.__init__ (data_wrangling/dataset.py:280-301): keep_small=False, seed=42
        self.sizes = [(int(w), int(h)) for w, h in sizes]
        self.batch_size, self.keep_small, self.shuffle, self.seed = batch_size, keep_small, shuffle, seed

    def __iter__(self) -> Iterator[List[int]]:
        batches = bucket_batches(self.sizes, self.batch_size, self.keep_small, self.shuffle, self.seed)
        if self.shuffle:
            self.seed += 1
        return iter(batches)

    def __len__(self) -> int:
        return len(bucket_batches(self.sizes, self.batch_size, self.keep_small, False))


def to_model_input(model, images, pad_multiple: int = 16):
    """uint8 images ((H, W) or (H, W, 1|3), numpy or torch, host or device) -> list of float32 (1, Hp, Wp) tensors on the
    model's device: ToTensor -> Grayscale(1) -> Invert, zero-padded to the encoder's 16-pixel grid."""
    return model.engine().preprocess_u8(images, pad_multiple)

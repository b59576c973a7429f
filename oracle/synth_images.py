"""TEST INFRASTRUCTURE ONLY -- synthetic image dataset + one feature's cache entries, shared by `gen_golden.py` (which
feeds them to the reference's image constructors) and the tests (which feed them to the mirror's)."""
from __future__ import annotations

import numpy as np
import torch


class FakeImageDataset:
    """Just enough of a HF `datasets.Dataset` for the image constructors: `len`, `.features`, `.select(indices=)[col]`."""

    def __init__(self, ids, images):
        self.ids, self.images = list(ids), list(images)
        self.features = {"id": None, "image": None}

    def __len__(self):
        return len(self.images)

    def select(self, indices):
        idx = [int(i) for i in (indices.tolist() if hasattr(indices, "tolist") else indices)]
        return {"id": [self.ids[i] for i in idx], "image": [self.images[i] for i in idx]}


def synth_image_cache(seed=51, n_images=80):
    """Synthetic image dataset (smooth gradients, every third image repeats the previous dataset id) and one feature's
    cache entries: (row, position) locations over positions 0..699, so that some fall outside the 576 base tokens."""
    from PIL import Image

    g = torch.Generator().manual_seed(seed)
    ids, images = [], []
    for i in range(n_images):
        ids.append(i - 1 if i % 3 == 2 else i)
        w, h = 40 + (i % 5) * 8, 30 + (i % 7) * 6
        yy, xx = np.mgrid[0:h, 0:w]
        arr = np.stack([(xx * 255 // w), (yy * 255 // h), np.full_like(xx, (i * 37) % 256)], -1).astype(np.uint8)
        images.append(Image.fromarray(arr, mode="RGB"))
    rows, pos, act = [], [], []
    for i in range(n_images):
        n = int(torch.randint(0, 60, (1,), generator=g))
        p = torch.randperm(700, generator=g)[:n]
        rows.append(torch.full((n,), i, dtype=torch.long))
        pos.append(p)
        act.append(torch.rand(n, generator=g) * (1 + (i % 11)))
    loc = torch.stack([torch.cat(rows), torch.cat(pos)], 1)
    return FakeImageDataset(ids, images), loc, torch.cat(act)

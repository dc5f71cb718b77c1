"""numpy restatement of the reference tiling path (TEST INFRASTRUCTURE, see oracle/__init__.py).

ImageSlicer plan / split / merge follow lib/tiles.py:35-161, the fusion weight lib/tiles.py:6-27,
NormalizeImage lib/augmentations.py:452-460, the D4 TTA lib/augmentations.py:476-511, and the dataset layout
lib/common.py:59-76.  The border is produced from an explicit reflect-101 index map instead of
cv2.copyMakeBorder, so the oracle has no OpenCV dependency.
"""
import math

import numpy as np

INRIA_MEAN = [0.40273115, 0.45046371, 0.42960134]   # lib/datasets/Inria.py:34
INRIA_STD = [3.15086464, 3.29831641, 3.63201004]    # lib/datasets/Inria.py:35


def reflect101_index(p, n):
    """Source index of cv2.BORDER_REFLECT_101 for padded coordinate p over n samples (edge not repeated)."""
    if n == 1:
        return 0
    while p < 0 or p >= n:
        p = -p if p < 0 else 2 * (n - 1) - p
    return p


def pyramid_weight_loop(width, height):
    """lib/tiles.py:6-27 evaluated pixel by pixel (slow; use for small sizes)."""
    xc, yc = width * 0.5, height * 0.5
    dc = np.zeros((width, height))
    de = np.zeros((width, height))
    for i in range(width):
        for j in range(height):
            dc[i, j] = np.sqrt(np.square(i - xc + 0.5) + np.square(j - yc + 0.5))
            cand = [np.sqrt(np.square(i - 0 + 0.5) + np.square(0.5)),
                    np.sqrt(np.square(i - width + 0.5) + np.square(0.5)),
                    np.sqrt(np.square(0.5) + np.square(j - 0 + 0.5)),
                    np.sqrt(np.square(0.5) + np.square(j - height + 0.5))]
            de[i, j] = np.min(cand)
    ratio = np.divide(de, np.add(dc, de))
    alpha = (width * height) / np.sum(ratio)
    return alpha * ratio


def pyramid_weight(width, height):
    """Array form of pyramid_weight_loop (same float64 operations per element; bit-equal, see the golden test)."""
    xc, yc = width * 0.5, height * 0.5
    i = np.arange(width, dtype=np.float64).reshape(-1, 1)
    j = np.arange(height, dtype=np.float64).reshape(1, -1)
    dc = np.sqrt(np.square(i - xc + 0.5) + np.square(j - yc + 0.5))
    q = np.square(0.5)
    left = np.broadcast_to(np.sqrt(np.square(i - 0 + 0.5) + q), (width, height))
    right = np.broadcast_to(np.sqrt(np.square(i - width + 0.5) + q), (width, height))
    bottom = np.broadcast_to(np.sqrt(q + np.square(j - 0 + 0.5)), (width, height))
    top = np.broadcast_to(np.sqrt(q + np.square(j - height + 0.5)), (width, height))
    de = np.minimum(np.minimum(left, right), np.minimum(bottom, top))
    ratio = np.divide(de, np.add(dc, de))
    alpha = (width * height) / np.sum(ratio)
    return alpha * ratio


class SlicerOracle:
    """Margins, crop list, split and merge of lib/tiles.py:35-161."""

    def __init__(self, image_shape, tile_size, tile_step=0, image_margin=0, weight='mean'):
        self.image_height, self.image_width = image_shape[0], image_shape[1]
        self.tile_size, self.tile_step = tile_size, tile_step
        if weight not in ('mean', 'pyramid'):
            raise KeyError(weight)
        self.weight = weight
        if tile_step < 1 or tile_step > tile_size:
            raise ValueError()
        overlap = tile_size - tile_step
        if image_margin == 0:
            nw = max(1, math.ceil((self.image_width - overlap) / tile_step))
            nh = max(1, math.ceil((self.image_height - overlap) / tile_step))
            extra_w = tile_step * nw - (self.image_width - overlap)
            extra_h = tile_step * nh - (self.image_height - overlap)
            self.margin_left = extra_w // 2
            self.margin_right = extra_w - self.margin_left
            self.margin_top = extra_h // 2
            self.margin_bottom = extra_h - self.margin_top
        else:
            if (self.image_width - overlap + 2 * image_margin) % tile_step != 0:
                raise ValueError()
            if (self.image_height - overlap + 2 * image_margin) % tile_step != 0:
                raise ValueError()
            self.margin_left = self.margin_right = self.margin_top = self.margin_bottom = image_margin
        padded_h = self.image_height + self.margin_top + self.margin_bottom
        padded_w = self.image_width + self.margin_left + self.margin_right
        self.crops = [(x, y, tile_size, tile_size)
                      for y in range(0, padded_h - tile_size + 1, tile_step)
                      for x in range(0, padded_w - tile_size + 1, tile_step)]

    def _padded(self, image):
        rows = [reflect101_index(p - self.margin_top, self.image_height)
                for p in range(self.image_height + self.margin_top + self.margin_bottom)]
        cols = [reflect101_index(p - self.margin_left, self.image_width)
                for p in range(self.image_width + self.margin_left + self.margin_right)]
        return image[np.asarray(rows)][:, np.asarray(cols)]

    def split(self, image):
        assert image.shape[0] == self.image_height and image.shape[1] == self.image_width
        padded = self._padded(image)
        return [padded[y:y + th, x:x + tw].copy() for x, y, tw, th in self.crops]

    def cut_patch(self, image, index):
        x, y, tw, th = self.crops[index]
        return self._padded(image)[y:y + th, x:x + tw].copy()

    def fusion_weight(self):
        if self.weight == 'mean':
            return np.ones((self.tile_size, self.tile_size), dtype=np.float32)
        return pyramid_weight(self.tile_size, self.tile_size)

    def merge(self, tiles, dtype=np.float32):
        if len(tiles) != len(self.crops):
            raise ValueError
        channels = 1 if tiles[0].ndim == 2 else tiles[0].shape[2]
        shape = (self.image_height + self.margin_top + self.margin_bottom,
                 self.image_width + self.margin_left + self.margin_right, channels)
        acc = np.zeros(shape, dtype=np.float64)
        norm = np.zeros(shape, dtype=np.float64)
        w = np.dstack([self.fusion_weight()] * channels)
        for tile, (x, y, tw, th) in zip(tiles, self.crops):   # crop order matters for float64 rounding
            acc[y:y + th, x:x + tw] += tile * w
            norm[y:y + th, x:x + tw] += w
        norm = np.clip(norm, a_min=np.finfo(norm.dtype).eps, a_max=None)
        merged = np.divide(acc, norm).astype(dtype)
        return merged[self.margin_top:self.margin_top + self.image_height,
                      self.margin_left:self.margin_left + self.image_width]


def normalize_image(x, mean=INRIA_MEAN, std=INRIA_STD, scale=1. / 255.):
    """NormalizeImage.__call__ (lib/augmentations.py:452-460): float32 mean/std arrays, float64 result for u8 input."""
    m = np.array(mean, dtype=np.float32)
    s = np.array(std, dtype=np.float32)
    return (x * float(scale) - m) / s


def normalize_lut(mean=INRIA_MEAN, std=INRIA_STD, scale=1. / 255.):
    """float32 [C][256]: the reference pipeline's value for every u8 input (normalise in float64, then .float())."""
    c = len(mean)
    levels = np.repeat(np.arange(256, dtype=np.uint8)[:, None], c, axis=1)       # [256][C] as an H x C 'image'
    return np.ascontiguousarray(normalize_image(levels, mean, std, scale).astype(np.float32).T)


def d4_views(image):
    """The 8 views of tta_d4_aug for one image, in its order (lib/augmentations.py:476-491)."""
    r = [image, np.rot90(image, 1), np.rot90(image, 2), np.rot90(image, 3)]
    return r + [np.fliplr(v) for v in r]


def tta_d4_aug(images):
    out = []
    for im in images:
        out.extend(d4_views(im))
    return out


def tta_d4_deaug(preds):
    """Mean of the 8 inverse-transformed predictions, float32 sum in the reference's order (:494-511)."""
    assert len(preds) % 8 == 0
    out = []
    for k in range(0, len(preds), 8):
        p = preds[k:k + 8]
        s = p[0] + np.rot90(p[1], -1) + np.rot90(p[2], -2) + np.rot90(p[3], -3)
        s = s + np.fliplr(p[4])
        s = s + np.rot90(np.fliplr(p[5]), -1)
        s = s + np.rot90(np.fliplr(p[6]), -2)
        s = s + np.rot90(np.fliplr(p[7]), -3)
        out.append(s * float(1. / 8.))
    return out


def to_nchw_float(tiles):
    """InMemoryDataset + DataLoader batch: list of HWC -> float32 [n][C][H][W] (lib/common.py:59-76)."""
    return np.stack([np.moveaxis(t, -1, 0) for t in tiles]).astype(np.float32)

"""The three pieces of reference lib/augmentations.py that sit on the inference path (:10-54, :452-460, :476-511).

`NormalizeImage` keeps the reference's call semantics on host arrays and additionally exposes the exact 256-entry
table the device split kernel applies to uint8 pixels (the value depends only on (level, channel), so evaluating
the reference's float64 expression once per level and rounding to float32 reproduces `.float()` of
lib/common.py:70 bit for bit).  The D4 TTA helpers work on CUDA tensors; the fused pipeline does not call them
(the split / merge kernels apply the same index maps), they exist for API parity.
"""
import numpy as np
import torch


class Sequential:
    def __init__(self, transforms):
        self.transforms = transforms

    def __call__(self, x, mask=None):
        for t in self.transforms:
            x, mask = t(x, mask)
        return x, mask


class ImageOnly:
    def __init__(self, trans):
        self.trans = trans

    def __call__(self, x, mask=None):
        return self.trans(x), mask


class NormalizeImage:
    def __init__(self, scale=1. / 255., mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225]):
        self.scale = float(scale)
        self.mean = np.array(mean, dtype=np.float32)
        self.std = np.array(std, dtype=np.float32)

    def __call__(self, x):
        x = (x * self.scale - self.mean) / self.std
        return x

    def lut(self):
        """float32 [C][256]: self(level) for every uint8 level and channel, rounded like `.float()`."""
        c = len(self.mean)
        levels = np.repeat(np.arange(256, dtype=np.uint8)[:, None], c, axis=1)
        return np.ascontiguousarray(self(levels).astype(np.float32).T)


def find_normalize(transform):
    """The NormalizeImage inside a reference-style test transform (Sequential([ImageOnly(NormalizeImage)]))."""
    if isinstance(transform, NormalizeImage):
        return transform
    if isinstance(transform, ImageOnly):
        return find_normalize(transform.trans)
    if isinstance(transform, Sequential):
        found = [find_normalize(t) for t in transform.transforms]
        found = [f for f in found if f is not None]
        if len(found) == 1 and len(transform.transforms) == 1:
            return found[0]
    return None


def _rot90(t, k):
    return torch.rot90(t, k, dims=(0, 1))


def tta_d4_aug(images):
    """8 dihedral views per H x W x C CUDA tensor, in the reference's order."""
    res = []
    for image in images:
        r = [image, _rot90(image, 1), _rot90(image, 2), _rot90(image, 3)]
        res.extend(r + [torch.flip(v, dims=(1,)) for v in r])
    return res


def tta_d4_deaug(image_list):
    """Mean of the 8 inverse-transformed predictions (float32 sum in the reference's order, then * 1/8)."""
    assert len(image_list) % 8 == 0
    res = []
    for i in range(0, len(image_list), 8):
        p = image_list[i:i + 8]
        fl = lambda v: torch.flip(v, dims=(1,))
        img = p[0] + _rot90(p[1], -1)
        img = img + _rot90(p[2], -2)
        img = img + _rot90(p[3], -3)
        img = img + fl(p[4])
        img = img + _rot90(fl(p[5]), -1)
        img = img + _rot90(fl(p[6]), -2)
        img = img + _rot90(fl(p[7]), -3)
        res.append(img * float(1. / 8.))
    return res

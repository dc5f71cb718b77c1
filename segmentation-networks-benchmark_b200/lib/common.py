"""Datasets of the hot path's input side (reference lib/common.py:41-159) with the tiling kept on the device.

`TiledImageDataset` is the training-side consumer of `ImageSlicer.cut_patch` (lib/common.py:116-159, used by
lib/datasets/Inria.py:64-65): the reference re-pads the whole 5000 x 5000 image on the host for EVERY patch it cuts.  Here
the image and its mask are uploaded once; a patch -- or a whole batch of consecutive patches -- is gathered by the split
kernels straight out of HBM (reflect-101 border included), and when the transform is the reference's
`ImageOnly(NormalizeImage(...))` the normalisation and the HWC -> CHW `.float()` of `__getitem__` are fused into the same
launch (snb_split_norm_u8, LAYOUT_NCHW_F32), so a sample never exists on the host.  Random augmentations stay out of scope
(SURVEY 2): any other transform is applied on the host exactly as the reference does, after a device -> host copy.
"""
import os

import numpy as np
import torch
from torch.utils.data import ConcatDataset, Dataset

from .. import _native as N
from .augmentations import find_normalize
from .tiles import ImageSlicer


def find_in_dir(dirname):
    return [os.path.join(dirname, fname) for fname in os.listdir(dirname)]


def read_rgb(fname):
    import cv2

    return cv2.imread(fname, cv2.IMREAD_COLOR)


def read_mask(fname):
    import cv2

    return cv2.imread(fname, cv2.IMREAD_GRAYSCALE)


def count_parameters(model):
    total = sum(p.numel() for p in model.parameters())
    trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
    return total, trainable


class InMemoryDataset(Dataset):
    """lib/common.py:52-76: (image HWC, mask) pairs -> (float CHW, long 1HW); kept for the reference call shape."""

    def __init__(self, images, masks, transform=None):
        self.images, self.masks, self.transform = images, masks, transform

    def __len__(self):
        return len(self.images)

    def __getitem__(self, index):
        i = self.images[index].copy()
        m = self.masks[index].copy() if self.masks is not None else None
        if self.transform is not None:
            i, m = self.transform(i, m)
        i = torch.from_numpy(np.moveaxis(i, -1, 0).copy()).float()
        if m is None:
            return i
        return i, torch.from_numpy(np.expand_dims(m, 0)).long()


class TiledImageDataset(Dataset):
    """lib/common.py:116-159 with the image, the mask and the tiling on the device.

    Same constructor as the reference plus `image=` / `mask=` (arrays instead of file names) and `device=`.  Items are
    (float [C, T, T], long [1, T, T]) CUDA tensors; `batch(begin, count)` returns `count` consecutive patches from ONE
    launch per tensor.  keep_in_mem is implied: the pair is uploaded once (75 + 25 MB for an Inria image)."""

    def __init__(self, image_fname=None, mask_fname=None, tile_size=512, tile_step=0, image_margin=0, transform=None,
                 target_shape=None, keep_in_mem=True, image=None, mask=None, device=None):
        N.require_cuda()
        self.image_fname, self.mask_fname = image_fname, mask_fname
        if image is None:
            image = read_rgb(image_fname)
        if mask is None:
            mask = read_mask(mask_fname)
        if image is None or mask is None:
            raise FileNotFoundError("could not read %r / %r" % (image_fname, mask_fname))
        if image.shape[0] != mask.shape[0] or image.shape[1] != mask.shape[1]:
            raise ValueError()                               # lib/common.py:129-130
        if target_shape is not None and tuple(target_shape[:2]) != tuple(image.shape[:2]):
            raise ValueError("target_shape %s does not match the image %s" % (tuple(target_shape), image.shape))
        if tile_step <= 0:
            tile_step = tile_size // 2
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.slicer = ImageSlicer(image.shape, tile_size, tile_step, image_margin)
        self.transform = transform
        self.norm = find_normalize(transform) if transform is not None else None
        if transform is not None and self.norm is None:
            self._host_transform = transform
        else:
            self._host_transform = None
        self.image = torch.from_numpy(np.ascontiguousarray(image)).to(self.device)
        self.mask = torch.from_numpy(np.ascontiguousarray(mask)).to(self.device)
        if self.image.dim() == 2:
            self.image = self.image.unsqueeze(-1)
        self.lut = torch.from_numpy(self.norm.lut()).to(self.device) if self.norm is not None else None

    def __len__(self):
        return len(self.slicer.crops)

    def batch(self, begin, count):
        """Patches [begin, begin + count) -> (float [count, C, T, T], long [count, 1, T, T]) on the device."""
        if not (0 <= begin and count >= 0 and begin + count <= len(self)):
            raise IndexError("patch range outside the %d crops" % len(self))
        T, s = self.slicer.tile_size, self.slicer
        c = self.image.shape[2]
        with torch.cuda.device(self.device):
            st = N.stream_ptr()
            if self._host_transform is None and self.image.dtype == torch.uint8 and self.lut is not None:
                # cut + NormalizeImage + HWC -> CHW .float() in one launch
                x = torch.empty((count, c, T, T), dtype=torch.float32, device=self.device)
                N.check(N.lib().snb_split_norm_u8(s.handle, N.ptr(self.image), c, N.ptr(self.lut), 0, N.LAYOUT_NCHW_F32, N.ptr(x),
                                                  begin, count, st))
            else:
                tiles = torch.empty((count, T, T, c), dtype=self.image.dtype, device=self.device)
                N.check(N.lib().snb_split_hwc(s.handle, N.ptr(self.image), c, self.image.element_size(), 0, None, N.ptr(tiles),
                                              begin, count, st))
                x = tiles
            m = torch.empty((count, T, T, 1), dtype=self.mask.dtype, device=self.device)
            N.check(N.lib().snb_split_hwc(s.handle, N.ptr(self.mask), 1, self.mask.element_size(), 0, None, N.ptr(m), begin, count, st))
        if self._host_transform is not None:
            # arbitrary (augmenting) transforms run on the host as in the reference, patch by patch
            xs, ms = [], []
            xh, mh = x.cpu().numpy(), m[..., 0].cpu().numpy()
            for i in range(count):
                xi, mi = self._host_transform(xh[i].copy(), mh[i].copy())
                xs.append(torch.from_numpy(np.moveaxis(xi, -1, 0).copy()).float())
                ms.append(torch.from_numpy(np.expand_dims(mi, 0)).long())
            return torch.stack(xs).to(self.device), torch.stack(ms).to(self.device)
        if x.dim() == 4 and x.shape[1] != c:                 # raw HWC tiles (no transform): moveaxis + .float()
            x = x.permute(0, 3, 1, 2).float()
        return x, m.permute(0, 3, 1, 2).long()

    def __getitem__(self, index):
        index = range(len(self))[index]
        x, m = self.batch(index, 1)
        return x[0], m[0]


class TiledImagesDataset(ConcatDataset):
    """lib/common.py:162-176: one TiledImageDataset per (image, mask) file pair."""

    def __init__(self, image_filenames, target_filenames, tile_size, tile_step=0, image_margin=0, target_shape=None,
                 transform=None, keep_in_mem=True, device=None):
        if len(image_filenames) != len(target_filenames):
            raise ValueError('Number of images does not corresponds to number of targets')
        super().__init__([TiledImageDataset(i, t, tile_size, tile_step, image_margin, transform, target_shape, keep_in_mem,
                                            device=device) for i, t in zip(image_filenames, target_filenames)])


def cut_dataset_in_patches(images, masks, patch_size, device=None):
    """lib/datasets/Inria.py:108-130 without the file I/O: every image / mask pair split into patch_size tiles with step
    patch_size // 2 on the device.  Yields (index, image tiles [n, T, T, C], mask tiles [n, T, T]) per pair."""
    slicer = None
    for k, (image, mask) in enumerate(zip(images, masks)):
        if slicer is None or (slicer.image_height, slicer.image_width) != tuple(image.shape[:2]):
            slicer = ImageSlicer(image.shape[:2], patch_size, patch_size // 2)
        ti = torch.stack(slicer.split(torch.from_numpy(np.ascontiguousarray(image)).cuda()))
        tm = torch.stack(slicer.split(torch.from_numpy(np.ascontiguousarray(mask)).cuda()))
        yield k, ti, tm

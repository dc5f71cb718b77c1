"""Drop-in `ImageSlicer` (reference lib/tiles.py:30-168) whose split / merge run as CUDA kernels.

Same constructor, attributes (`image_height/width, tile_size, tile_step, margin_*`, `crops`), methods and error
behaviour as the reference class.  Inputs may be numpy arrays (copied to the current CUDA device and back, so the
call sites of inria_submit.py:240-256 work unchanged) or CUDA tensors (results stay on the device).
The crop plan lives in the native library (snb_slicer_*); the pyramid weight is evaluated once per tile size on
the host with the reference's exact float64 expression (lib/tiles.py:6-27) and cached on the device.
"""
import ctypes

import numpy as np
import torch

from .. import _native as N

BORDER_CONSTANT = 0      # cv2.BORDER_CONSTANT
BORDER_REFLECT101 = 4    # cv2.BORDER_REFLECT_101 (== BORDER_DEFAULT)

_NP_DT = {np.dtype(np.uint8): N.DT_U8, np.dtype(np.float32): N.DT_F32, np.dtype(np.float64): N.DT_F64}
_TORCH_DT = {torch.uint8: N.DT_U8, torch.float32: N.DT_F32, torch.float64: N.DT_F64}


def compute_patch_weight_loss(width, height):
    """Vectorised, bit-equal restatement of lib/tiles.py:6-27 (same float64 operations element by element)."""
    xc = width * 0.5
    yc = height * 0.5
    xl, xr, yb, yt = 0, width, 0, height
    i = np.arange(width, dtype=np.float64)[:, None]
    j = np.arange(height, dtype=np.float64)[None, :]
    zero_half = 0.5  # (j - j + 0.5) and (i - i + 0.5) in the reference
    Dc = np.sqrt(np.square(i - xc + 0.5) + np.square(j - yc + 0.5))
    De_l = np.sqrt(np.square(i - xl + 0.5) + np.square(zero_half)) + 0 * j
    De_r = np.sqrt(np.square(i - xr + 0.5) + np.square(zero_half)) + 0 * j
    De_b = np.sqrt(np.square(zero_half) + np.square(j - yb + 0.5)) + 0 * i
    De_t = np.sqrt(np.square(zero_half) + np.square(j - yt + 0.5)) + 0 * i
    De = np.minimum(np.minimum(De_l, De_r), np.minimum(De_b, De_t))
    ratio = np.divide(De, np.add(Dc, De))
    alpha = (width * height) / np.sum(ratio)
    W = alpha * ratio
    return W, Dc, De


class ImageSlicer:
    """Helper class to slice image into tiles and merge them back with fusion (reference lib/tiles.py:30)."""

    def __init__(self, image_shape, tile_size, tile_step=0, image_margin=0, weight='mean'):
        self.image_height = image_shape[0]
        self.image_width = image_shape[1]
        self.tile_size = tile_size
        self.tile_step = tile_step

        weights = {'mean': self._mean, 'pyramid': self._pyramid}
        self.compute_weight = weights[weight]  # KeyError on unknown names, as in the reference
        self.weight = weight

        handle = ctypes.c_void_p()
        # tile_step < 1 / > tile_size and non-tiling image_margin raise ValueError (lib/tiles.py:56-57,81-85)
        N.check(N.lib().snb_slicer_create(int(self.image_height), int(self.image_width), int(tile_size),
                                          int(tile_step), int(image_margin), ctypes.byref(handle)))
        self._h = handle
        info = (ctypes.c_int64 * 8)()
        N.check(N.lib().snb_slicer_info(self._h, info))
        self.margin_left, self.margin_right, self.margin_top, self.margin_bottom = (int(v) for v in info[:4])
        n_tiles = int(info[4])
        self.tiles_x, self.tiles_y = int(info[5]), int(info[6])
        xy = (ctypes.c_int64 * (2 * max(n_tiles, 1)))()
        N.check(N.lib().snb_slicer_crops(self._h, xy))
        self.crops = [(int(xy[2 * i]), int(xy[2 * i + 1]), tile_size, tile_size) for i in range(n_tiles)]
        self._weight_dev = {}

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                N.lib().snb_slicer_destroy(h)
            except Exception:
                pass
            self._h = None

    # ------------------------------------------------------------------------------------------ helpers
    @property
    def handle(self):
        return self._h

    def _to_device(self, image):
        """-> (contiguous CUDA tensor H x W x C, was_numpy, had_channel_dim)"""
        N.require_cuda()
        was_numpy = isinstance(image, np.ndarray)
        t = torch.from_numpy(np.ascontiguousarray(image)).cuda() if was_numpy else image
        if not t.is_cuda:
            raise RuntimeError("ImageSlicer works on numpy arrays or CUDA tensors (no CPU fallback)")
        had_c = t.dim() == 3
        if not had_c:
            t = t.unsqueeze(-1)
        return t.contiguous(), was_numpy, had_c

    def weight_on_device(self, device=None):
        """float64 [T][T] fusion weight, cached per device."""
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        if dev not in self._weight_dev:
            w = np.asarray(self.compute_weight(self.tile_size), dtype=np.float64)
            self._weight_dev[dev] = torch.from_numpy(np.ascontiguousarray(w)).to(dev)
        return self._weight_dev[dev]

    def _border(self, t, borderType, value):
        if borderType == BORDER_REFLECT101:
            return 0, None
        if borderType == BORDER_CONSTANT:
            # cv2 semantics: a scalar fills channel 0 and leaves the others 0; a sequence fills per channel
            c = t.shape[2]
            vals = np.zeros(c, dtype=np.float64)
            seq = np.atleast_1d(np.asarray(value, dtype=np.float64))
            vals[:min(c, len(seq), 4)] = seq[:min(c, len(seq), 4)]
            np_dt = torch.empty(0, dtype=t.dtype).numpy().dtype
            if np.issubdtype(np_dt, np.integer):
                ii = np.iinfo(np_dt)
                vals = np.clip(np.rint(vals), ii.min, ii.max)  # cv2 saturate_cast
            buf = vals.astype(np_dt).tobytes()
            return 1, ctypes.create_string_buffer(buf, len(buf))
        raise NotImplementedError("borderType %r (only BORDER_REFLECT101 and BORDER_CONSTANT)" % (borderType,))

    def _split_range(self, image, begin, count, borderType, value):
        assert image.shape[0] == self.image_height
        assert image.shape[1] == self.image_width
        t, was_numpy, had_c = self._to_device(image)
        mode, border = self._border(t, borderType, value)
        c = t.shape[2]
        out = torch.empty((count, self.tile_size, self.tile_size, c), dtype=t.dtype, device=t.device)
        N.check(N.lib().snb_split_hwc(self._h, N.ptr(t), c, t.element_size(), mode,
                                      ctypes.cast(border, ctypes.c_void_p) if border is not None else None,
                                      N.ptr(out), begin, count, N.stream_ptr()))
        if not had_c:
            out = out[..., 0]
        return out, was_numpy

    # ---------------------------------------------------------------------------------------------- API
    def split(self, image, borderType=BORDER_REFLECT101, value=0):
        """List of T x T x C tile copies in crop order (2-D in -> 2-D tiles), lib/tiles.py:98-117."""
        out, was_numpy = self._split_range(image, 0, len(self.crops), borderType, value)
        if was_numpy:
            host = out.cpu().numpy()
            return [host[i] for i in range(host.shape[0])]
        return list(out.unbind(0))

    def cut_patch(self, image, slice_index, borderType=BORDER_REFLECT101, value=0):
        """Single crop, lib/tiles.py:119-135 (no full-image re-pad: the kernel gathers just this tile)."""
        index = range(len(self.crops))[slice_index]  # IndexError / negative indices like list indexing
        out, was_numpy = self._split_range(image, index, 1, borderType, value)
        return out[0].cpu().numpy() if was_numpy else out[0]

    def merge(self, tiles, dtype=np.float32):
        """Weighted overlap-add of per-tile arrays back to H x W x C, lib/tiles.py:137-161."""
        if len(tiles) != len(self.crops):
            raise ValueError
        N.require_cuda()
        was_numpy = isinstance(tiles[0], np.ndarray)
        if was_numpy:
            t = torch.from_numpy(np.ascontiguousarray(np.stack(tiles))).cuda()
        else:
            t = tiles if isinstance(tiles, torch.Tensor) else torch.stack(list(tiles))
        if t.dim() != 4:
            # the reference broadcasts (T,T) * (T,T,1) and fails on the slice assignment for 2-D tiles
            raise ValueError("merge expects T x T x C tiles")
        if t.dtype not in _TORCH_DT:
            t = t.double()
        t = t.contiguous()
        out_dt = np.dtype(dtype)
        if out_dt not in _NP_DT:
            raise NotImplementedError("merge dtype %s" % out_dt)
        tdt = {N.DT_U8: torch.uint8, N.DT_F32: torch.float32, N.DT_F64: torch.float64}[_NP_DT[out_dt]]
        c = t.shape[3]
        out = torch.empty((self.image_height, self.image_width, c), dtype=tdt, device=t.device)
        w = self.weight_on_device(t.device)
        N.check(N.lib().snb_merge(self._h, N.ptr(t), _TORCH_DT[t.dtype], c, 1, N.ptr(w), N.ptr(out),
                                  _NP_DT[out_dt], None, 0.5, N.stream_ptr()))
        return out.cpu().numpy() if was_numpy else out

    def _mean(self, tile_size):
        return np.ones((tile_size, tile_size), dtype=np.float32)

    def _pyramid(self, tile_size):
        w, _, _ = compute_patch_weight_loss(tile_size, tile_size)
        return w

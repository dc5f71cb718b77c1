"""ZF_UNET (reference lib/models/zf_unet.py:35-95): UNet with BatchNorm, nearest x2 upsampling and concat skips.

Same constructor, module tree and state_dict keys (156 entries with batch_norm=True); the forward pass runs on the
native sm_100a engine in eval mode (BatchNorm folded into the convolutions, Dropout2d inactive, lib/models/zf_unet.py:9,26).
Training-mode forward (batch statistics, dropout masks) is not on the inference path and raises.
"""
import torch
from torch import nn

from ... import _native as N
from ...engine import PRECISIONS, ZFUNetPlan


class _Conv3BN(nn.Module):
    def __init__(self, in_: int, out: int, bn=False):
        super().__init__()
        self.conv = nn.Conv2d(in_, out, 3, padding=1)
        self.bn = nn.BatchNorm2d(out) if bn else None
        self.activation = nn.ReLU(inplace=True)

    def folded(self):
        bn = None
        if self.bn is not None:
            bn = (self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var, self.bn.eps)
        return self.conv.weight, self.conv.bias, bn


class _DoubleConvModule(nn.Module):
    def __init__(self, in_: int, out: int, dropout_val, batch_norm):
        super().__init__()
        self.l1 = _Conv3BN(in_, out, batch_norm)
        self.l2 = _Conv3BN(out, out, batch_norm)
        self.dropout = nn.Dropout2d(p=dropout_val)


class ZF_UNET(nn.Module):
    def __init__(self, dropout_val=0.2, batch_norm=True, input_channels=3, num_classes=1, filters=32):
        super(ZF_UNET, self).__init__()
        self.num_classes = num_classes
        self.pool = nn.MaxPool2d(2)
        self.unpool = nn.Upsample(scale_factor=2)
        f = filters
        # (attribute name, input channels, output channels): encoder widths double per level, the decoder blocks
        # consume [upsampled deeper features | skip] (lib/models/zf_unet.py:44-58 keeps these exact names)
        table = [('conv_224', input_channels, f), ('conv_112', f, 2 * f), ('conv_56', 2 * f, 4 * f),
                 ('conv_28', 4 * f, 8 * f), ('conv_14', 8 * f, 16 * f), ('conv_7', 16 * f, 32 * f)]
        for level, name in enumerate(['up_conv_14', 'up_conv_28', 'up_conv_56', 'up_conv_112', 'up_conv_224']):
            wide = (32 * f) >> level                      # channels arriving through the x2 upsampling
            table.append((name, wide + wide // 2, wide // 2))
        for name, c_in, c_out in table:
            setattr(self, name, _DoubleConvModule(c_in, c_out, dropout_val, batch_norm))
        self.conv_final = nn.Conv2d(f, num_classes, 1)

    precision = 'bf16'    # or 'tf32': fp32 storage + TF32 tensor-core products (probabilities within 1e-4)

    def set_precision(self, precision):
        if precision not in PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(PRECISIONS))
        self.precision = precision
        return self

    def _blocks(self):
        return [self.conv_224, self.conv_112, self.conv_56, self.conv_28, self.conv_14, self.conv_7, self.up_conv_14,
                self.up_conv_28, self.up_conv_56, self.up_conv_112, self.up_conv_224]

    def _stamp(self):
        tensors = list(self.parameters()) + list(self.buffers())
        return tuple((t.data_ptr(), t._version) for t in tensors)

    def plan(self, n, h, w, sigmoid=False):
        """Cached ZFUNetPlan; rebuilt when parameters or BatchNorm buffers were modified or moved."""
        cache = self.__dict__.setdefault('_plans', {})
        stamp = self._stamp()
        if self.__dict__.get('_plan_stamp') != stamp:
            cache.clear()
            self.__dict__['_plan_stamp'] = stamp
        precision = getattr(self, 'precision', 'bf16')
        if precision not in PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(PRECISIONS))
        key = (n, h, w, bool(sigmoid), precision)
        if key not in cache:
            dev = self.conv_final.weight.device
            if dev.type != 'cuda':
                raise RuntimeError("ZF_UNET runs on CUDA devices only (no CPU fallback); call .cuda()")
            blocks = [(b.l1.folded(), b.l2.folded()) for b in self._blocks()]
            with torch.no_grad():
                cache[key] = ZFUNetPlan(blocks, (self.conv_final.weight, self.conv_final.bias), n, h, w, dev, sigmoid,
                                        PRECISIONS[precision])
        return cache[key]

    def forward(self, x):
        N.require_cuda()
        if self.training:
            raise NotImplementedError("ZF_UNET on the native engine is inference only: call .eval() "
                                      "(training-mode BatchNorm / Dropout2d are not on the tiled-inference path)")
        if not x.is_cuda:
            raise RuntimeError("input must be a CUDA tensor (no CPU fallback)")
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("expected input of shape [N, 3, H, W]")
        with torch.cuda.device(x.device):
            p = self.plan(x.shape[0], x.shape[2], x.shape[3], sigmoid=False)
            p.load_nchw(x.float())
            out = p.run()
        return out.unsqueeze(1).clone()

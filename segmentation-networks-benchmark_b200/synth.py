"""Deterministic synthetic weights and inputs (state_dicts with the reference's key names, Inria-shaped images, logits /
targets) shared by bench.py, the tools, the golden generator and the tests.  Pure data generation from numpy RandomState
seeds -- no algorithm of the hot path lives here -- so the GPU box regenerates bit-identical tensors without any file
transfer.  It lives in the package (not under oracle/) so that the CUDA arm of bench.py never imports the oracle."""
import numpy as np
import torch

# (out_channels of each conv, in torchvision `features` index order); 'M' = max-pool
_VGG = {
    'unet16': [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M'],
    'unet11': [64, 'M', 128, 'M', 256, 256, 'M', 512, 512, 'M', 512, 512, 'M'],
}
# decoder blocks (in, mid, out) for num_filters = 32: center, dec5, dec4, dec3, dec2
_DEC = {
    'unet16': [(512, 512, 256), (768, 512, 256), (768, 512, 256), (512, 256, 64), (192, 128, 32)],
    'unet11': [(512, 512, 256), (768, 512, 256), (768, 512, 128), (384, 256, 64), (192, 128, 32)],
}
_ALIASES = {   # convK.M -> encoder.N, the double registration of lib/models/unet16.py:73-102 / unet11.py:71-93
    'unet16': {'conv1': [0, 2], 'conv2': [5, 7], 'conv3': [10, 12, 14], 'conv4': [17, 19, 21], 'conv5': [24, 26, 28]},
    'unet11': {'conv1': [0], 'conv2': [3], 'conv3': [6, 8], 'conv4': [11, 13], 'conv5': [16, 18]},
}


def _he(rs, shape, fan_in, gain=1.0):
    return torch.from_numpy((rs.standard_normal(shape) * gain * np.sqrt(2.0 / fan_in)).astype(np.float32))


def _bias(rs, n):
    return torch.from_numpy((rs.standard_normal(n) * 0.05).astype(np.float32))


def vgg_unet_state_dict(arch='unet16', seed=0):
    """Random (He-scaled) state_dict with the reference's key names and shapes, aliases included."""
    rs = np.random.RandomState(seed)
    sd = {}
    cin, idx = 3, 0
    for v in _VGG[arch]:
        if v == 'M':
            idx += 1
            continue
        sd['encoder.%d.weight' % idx] = _he(rs, (v, cin, 3, 3), cin * 9)
        sd['encoder.%d.bias' % idx] = _bias(rs, v)
        cin = v
        idx += 2
    for name, positions in _ALIASES[arch].items():
        for k, enc_idx in enumerate(positions):
            for leaf in ('weight', 'bias'):
                sd['%s.%d.%s' % (name, 2 * k, leaf)] = sd['encoder.%d.%s' % (enc_idx, leaf)]
    for name, (cin_d, mid, out) in zip(['center', 'dec5', 'dec4', 'dec3', 'dec2'], _DEC[arch]):
        sd[name + '.block.0.conv.weight'] = _he(rs, (mid, cin_d, 3, 3), cin_d * 9)
        sd[name + '.block.0.conv.bias'] = _bias(rs, mid)
        sd[name + '.block.1.weight'] = _he(rs, (mid, out, 4, 4), mid * 4)   # 4 taps reach each output pixel
        sd[name + '.block.1.bias'] = _bias(rs, out)
    sd['dec1.conv.weight'] = _he(rs, (32, 96, 3, 3), 96 * 9)
    sd['dec1.conv.bias'] = _bias(rs, 32)
    sd['final.weight'] = _he(rs, (1, 32, 1, 1), 32, gain=0.5)   # logits ~N(0,1): probabilities span (0,1)
    sd['final.bias'] = _bias(rs, 1)
    return sd


def zf_unet_state_dict(seed=0, filters=32):
    """Random ZF_UNET state_dict (reference key names); BatchNorm buffers randomised so the folding is exercised
    (running_mean~N(0,0.1), running_var~U(0.5,1.5), weight~U(0.5,1.5), bias~N(0,0.1))."""
    rs = np.random.RandomState(seed)
    f = filters
    io = [('conv_224', 3, f), ('conv_112', f, 2 * f), ('conv_56', 2 * f, 4 * f), ('conv_28', 4 * f, 8 * f),
          ('conv_14', 8 * f, 16 * f), ('conv_7', 16 * f, 32 * f), ('up_conv_14', 48 * f, 16 * f),
          ('up_conv_28', 24 * f, 8 * f), ('up_conv_56', 12 * f, 4 * f), ('up_conv_112', 6 * f, 2 * f),
          ('up_conv_224', 3 * f, f)]
    sd = {}
    for name, cin, cout in io:
        for layer, ci in (('l1', cin), ('l2', cout)):
            pre = '%s.%s.' % (name, layer)
            # gain 0.8: with the randomised BatchNorm scales (mean square ~1.2) activations keep O(1) magnitude over the
            # 22 layers, so the bf16 rounding noise of the logits stays ~1e-2 (probabilities inside the 2e-2 band)
            sd[pre + 'conv.weight'] = _he(rs, (cout, ci, 3, 3), ci * 9, gain=0.8)
            sd[pre + 'conv.bias'] = _bias(rs, cout)
            sd[pre + 'bn.weight'] = torch.from_numpy(rs.uniform(0.5, 1.5, cout).astype(np.float32))
            sd[pre + 'bn.bias'] = torch.from_numpy((rs.standard_normal(cout) * 0.1).astype(np.float32))
            sd[pre + 'bn.running_mean'] = torch.from_numpy((rs.standard_normal(cout) * 0.1).astype(np.float32))
            sd[pre + 'bn.running_var'] = torch.from_numpy(rs.uniform(0.5, 1.5, cout).astype(np.float32))
            sd[pre + 'bn.num_batches_tracked'] = torch.tensor(7, dtype=torch.int64)
    sd['conv_final.weight'] = _he(rs, (1, f, 1, 1), f, gain=1.5)
    sd['conv_final.bias'] = _bias(rs, 1)
    return sd


def fcdensenet_state_dict(seed=0, down_blocks=(5, 5, 5, 5, 5), up_blocks=(5, 5, 5, 5, 5), bottleneck_layers=5, growth=16,
                          first=48, n_classes=1):
    """Random FCDenseNet state_dict with the reference key names (434 entries for FCDenseNet67) and randomised BatchNorm
    buffers."""
    rs = np.random.RandomState(seed)
    sd = {}

    def bn(prefix, c):
        sd[prefix + '.weight'] = torch.from_numpy(rs.uniform(0.5, 1.5, c).astype(np.float32))
        sd[prefix + '.bias'] = torch.from_numpy((rs.standard_normal(c) * 0.1).astype(np.float32))
        sd[prefix + '.running_mean'] = torch.from_numpy((rs.standard_normal(c) * 0.1).astype(np.float32))
        sd[prefix + '.running_var'] = torch.from_numpy(rs.uniform(0.5, 1.5, c).astype(np.float32))
        sd[prefix + '.num_batches_tracked'] = torch.tensor(3, dtype=torch.int64)

    def dense_block(prefix, cin, n_layers):
        for k in range(n_layers):
            c = cin + k * growth
            bn('%s.layers.%d.norm' % (prefix, k), c)
            sd['%s.layers.%d.conv.weight' % (prefix, k)] = _he(rs, (growth, c, 3, 3), c * 9)
            sd['%s.layers.%d.conv.bias' % (prefix, k)] = _bias(rs, growth)

    sd['firstconv.weight'] = _he(rs, (first, 3, 3, 3), 27)
    sd['firstconv.bias'] = _bias(rs, first)
    cur, skips = first, []
    for i, n_layers in enumerate(down_blocks):
        dense_block('denseBlocksDown.%d' % i, cur, n_layers)
        cur += growth * n_layers
        skips.insert(0, cur)
        bn('transDownBlocks.%d.norm' % i, cur)
        sd['transDownBlocks.%d.conv.weight' % i] = _he(rs, (cur, cur, 1, 1), cur)
        sd['transDownBlocks.%d.conv.bias' % i] = _bias(rs, cur)
    dense_block('bottleneck.bottleneck', cur, bottleneck_layers)
    prev = growth * bottleneck_layers
    for i, n_layers in enumerate(up_blocks):
        sd['transUpBlocks.%d.convTrans.weight' % i] = _he(rs, (prev, prev, 3, 3), prev * 2.25)
        sd['transUpBlocks.%d.convTrans.bias' % i] = _bias(rs, prev)
        cur = prev + skips[i]
        dense_block('denseBlocksUp.%d' % i, cur, n_layers)
        prev = growth * n_layers
        cur += prev
    sd['finalConv.weight'] = _he(rs, (n_classes, cur, 1, 1), cur, gain=0.1)   # logits ~N(0,1) over the 288-wide concat
    sd['finalConv.bias'] = _bias(rs, n_classes)
    return sd


def linknet34_state_dict(seed=0):
    """Random LinkNet34 state_dict with the reference key names (294 entries); BatchNorm / InPlaceABN buffers randomised,
    some InPlaceABN weights negative so that the |weight| + eps scale of the backend matters."""
    rs = np.random.RandomState(seed)
    sd = {}

    def bn(prefix, c, tracked=True, signed=False):
        wgt = rs.uniform(0.5, 1.5, c).astype(np.float32)
        if signed:
            wgt *= np.where(rs.rand(c) < 0.25, -1.0, 1.0).astype(np.float32)
        sd[prefix + '.weight'] = torch.from_numpy(wgt)
        sd[prefix + '.bias'] = torch.from_numpy((rs.standard_normal(c) * 0.1).astype(np.float32))
        sd[prefix + '.running_mean'] = torch.from_numpy((rs.standard_normal(c) * 0.1).astype(np.float32))
        sd[prefix + '.running_var'] = torch.from_numpy(rs.uniform(0.5, 1.5, c).astype(np.float32))
        if tracked:
            sd[prefix + '.num_batches_tracked'] = torch.tensor(5, dtype=torch.int64)

    sd['firstconv.weight'] = _he(rs, (64, 3, 7, 7), 147)
    bn('firstbn', 64)
    filters, inplanes = [64, 128, 256, 512], 64
    for li, (planes, blocks) in enumerate(zip(filters, (3, 4, 6, 3))):
        for b in range(blocks):
            pre = 'encoder%d.%d' % (li + 1, b)
            cin = inplanes if b == 0 else planes
            sd[pre + '.conv1.weight'] = _he(rs, (planes, cin, 3, 3), cin * 9, gain=0.8)
            bn(pre + '.bn1', planes)
            sd[pre + '.conv2.weight'] = _he(rs, (planes, planes, 3, 3), planes * 9, gain=0.5)
            bn(pre + '.bn2', planes)
            if b == 0 and li > 0:
                sd[pre + '.downsample.0.weight'] = _he(rs, (planes, cin, 1, 1), cin, gain=0.7)
                bn(pre + '.downsample.1', planes)
        inplanes = planes
    for i in range(4, 0, -1):
        cin, cout = filters[i - 1], filters[max(i - 2, 0)]
        q, pre = cin // 4, 'decoder%d' % i
        sd[pre + '.conv1.weight'] = _he(rs, (q, cin, 1, 1), cin)
        sd[pre + '.conv1.bias'] = _bias(rs, q)
        bn(pre + '.abn1', q, tracked=False, signed=True)
        sd[pre + '.deconv2.weight'] = _he(rs, (q, q, 4, 4), q * 4)
        sd[pre + '.deconv2.bias'] = _bias(rs, q)
        bn(pre + '.abn2', q, tracked=False, signed=True)
        sd[pre + '.conv3.weight'] = _he(rs, (cout, q, 1, 1), q, gain=0.7)
        sd[pre + '.conv3.bias'] = _bias(rs, cout)
        bn(pre + '.abn3', cout, tracked=False, signed=True)
    sd['finaldeconv1.weight'] = _he(rs, (64, 32, 3, 3), 64 * 2.25)
    sd['finaldeconv1.bias'] = _bias(rs, 32)
    sd['finalconv2.weight'] = _he(rs, (32, 32, 3, 3), 32 * 9)
    sd['finalconv2.bias'] = _bias(rs, 32)
    sd['finalconv3.weight'] = _he(rs, (1, 32, 2, 2), 32 * 4, gain=0.12)   # logits within +-3: bf16 noise stays < 2e-2 in probability
    sd['finalconv3.bias'] = _bias(rs, 1)
    return sd


def image_u8(seed, h, w, c=3, smooth=True):
    """Inria-shaped synthetic uint8 image; low-pass structure so masks are not pure noise."""
    rs = np.random.RandomState(seed)
    if not smooth:
        return rs.randint(0, 256, (h, w, c)).astype(np.uint8)
    ch, cw = (h + 31) // 32 + 1, (w + 31) // 32 + 1
    coarse = rs.rand(ch, cw, c).astype(np.float32)
    up = np.repeat(np.repeat(coarse, 32, axis=0), 32, axis=1)[:h, :w]
    noise = rs.rand(h, w, c).astype(np.float32)
    return np.clip((0.7 * up + 0.3 * noise) * 255.0, 0, 255).astype(np.uint8)


def gt_mask_u8(seed, h, w):
    """Synthetic ground truth {0,1} (RandomState(1000+i).rand(h, w) > 0.5)."""
    return (np.random.RandomState(1000 + seed).rand(h, w) > 0.5).astype(np.uint8)


def logits_targets(seed, shape):
    rs = np.random.RandomState(seed)
    logits = torch.from_numpy(rs.standard_normal(shape).astype(np.float32))
    targets = torch.from_numpy((rs.rand(*shape) > 0.5).astype(np.int64))
    return logits, targets


def unet_state_dict(seed=0, abn=False, n_filters=32):
    """He-initialised weights with the reference's UNet / UNetABN key names (lib/models/unet.py, unet_abn.py); norm buffers
    randomised so that folding is exercised; ABN weights get mixed signs (the |weight| + eps scale)."""
    rs = np.random.RandomState(seed)
    sd = {}
    nf = n_filters

    def double_conv(prefix, cin, cout):
        i2 = 2 if abn else 3
        for ci, c_in in ((0, cin), (i2, cout)):
            sd['%s.%d.weight' % (prefix, ci)] = _he(rs, (cout, c_in, 3, 3), 9 * c_in, 0.9)
            sd['%s.%d.bias' % (prefix, ci)] = _bias(rs, cout)
            n = '%s.%d' % (prefix, ci + 1)
            w = torch.from_numpy(rs.uniform(0.5, 1.5, cout).astype(np.float32))
            if abn:
                w = w * torch.from_numpy(np.where(rs.rand(cout) < 0.3, -1.0, 1.0).astype(np.float32))
            sd[n + '.weight'] = w
            sd[n + '.bias'] = torch.from_numpy((rs.standard_normal(cout) * 0.1).astype(np.float32))
            sd[n + '.running_mean'] = torch.from_numpy((rs.standard_normal(cout) * 0.1).astype(np.float32))
            sd[n + '.running_var'] = torch.from_numpy(rs.uniform(0.5, 1.5, cout).astype(np.float32))
            if not abn:
                sd[n + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.int64)

    double_conv('inc.conv.conv', 3, nf)
    for i, (a, b) in enumerate([(nf, 2 * nf), (2 * nf, 4 * nf), (4 * nf, 8 * nf), (8 * nf, 8 * nf)], 1):
        double_conv('down%d.mpconv.1.conv' % i, a, b)
    for i, (a, b) in enumerate([(16 * nf, 4 * nf), (8 * nf, 2 * nf), (4 * nf, nf), (2 * nf, nf)], 1):
        double_conv('up%d.conv.conv' % i, a, b)
    sd['outc.conv.weight'] = _he(rs, (1, nf, 1, 1), nf, 0.5)
    sd['outc.conv.bias'] = _bias(rs, 1)
    return sd

"""CPU PyTorch fp32 restatement of the network / loss / metric part of the hot path (TEST INFRASTRUCTURE).

unet_vgg_forward follows lib/models/unet16.py:113-131 and lib/models/unet11.py:106-122 functionally from a
state_dict (no nn.Module, no torchvision); losses follow lib/losses.py:31-75, metrics lib/metrics.py:9-40 and
lib/train_utils.py:109-125.
"""
import numpy as np
import torch
import torch.nn.functional as F

VGG16_STAGES = [[0, 2], [5, 7], [10, 12, 14], [17, 19, 21], [24, 26, 28]]   # encoder.N indices per stage
VGG11_STAGES = [[0], [3], [6, 8], [11, 13], [16, 18]]
STAGES = {'unet16': VGG16_STAGES, 'unet11': VGG11_STAGES}


def _conv_relu(x, sd, prefix, quant=None):
    w, b = sd[prefix + '.weight'], sd[prefix + '.bias']
    if quant is not None:
        x, w = quant(x), quant(w)
    return F.relu(F.conv2d(x, w, b, padding=1))


def _decoder(x, sd, name, quant=None):
    x = _conv_relu(x, sd, name + '.block.0.conv', quant)
    w, b = sd[name + '.block.1.weight'], sd[name + '.block.1.bias']
    if quant is not None:
        x, w = quant(x), quant(w)
    return F.relu(F.conv_transpose2d(x, w, b, stride=2, padding=1))


def unet_vgg_forward(sd, x, arch='unet16', quant=None):
    """Logits [N,1,H,W].  `quant` (e.g. a bf16 round trip) is applied to every conv's input and weight to model
    the precision of the bf16 tensor-core path while keeping fp32 accumulation."""
    skips = []
    for stage in STAGES[arch]:
        for idx in stage:
            x = _conv_relu(x, sd, 'encoder.%d' % idx, quant)
        skips.append(x)
        x = F.max_pool2d(x, 2, 2)
    x = _decoder(x, sd, 'center', quant)
    for name, skip in zip(['dec5', 'dec4', 'dec3', 'dec2'], skips[:0:-1]):
        x = _decoder(torch.cat([x, skip], 1), sd, name, quant)
    x = _conv_relu(torch.cat([x, skips[0]], 1), sd, 'dec1.conv', quant)
    w, b = sd['final.weight'], sd['final.bias']
    return F.conv2d(x, w, b)   # the 1x1 head stays fp32 on the device path too


ZF_BLOCKS = ['conv_224', 'conv_112', 'conv_56', 'conv_28', 'conv_14', 'conv_7', 'up_conv_14', 'up_conv_28',
             'up_conv_56', 'up_conv_112', 'up_conv_224']


def _conv_bn_relu(x, sd, prefix, quant=None, fold=False):
    """_Conv3BN.forward in eval mode (lib/models/zf_unet.py:5-17).  With fold=True the BatchNorm is folded into the
    conv first (what the device path does), which matters only when `quant` rounds the folded weights."""
    w, b = sd[prefix + '.conv.weight'], sd[prefix + '.conv.bias']
    g, beta = sd[prefix + '.bn.weight'], sd[prefix + '.bn.bias']
    mean, var = sd[prefix + '.bn.running_mean'], sd[prefix + '.bn.running_var']
    if fold:
        scale = g / torch.sqrt(var + 1e-5)
        w, b = w * scale.view(-1, 1, 1, 1), (b - mean) * scale + beta
        if quant is not None:
            x, w = quant(x), quant(w)
        return F.relu(F.conv2d(x, w, b, padding=1))
    y = F.conv2d(x, w, b, padding=1)
    return F.relu(F.batch_norm(y, mean, var, g, beta, training=False, eps=1e-5))


def zf_unet_forward(sd, x, quant=None, fold=False):
    """ZF_UNET.forward in eval mode (lib/models/zf_unet.py:60-95): Dropout2d inactive, nearest x2 upsampling."""
    def block(t, name):
        t = _conv_bn_relu(t, sd, name + '.l1', quant, fold)
        return _conv_bn_relu(t, sd, name + '.l2', quant, fold)

    skips = []
    for name in ZF_BLOCKS[:5]:
        x = block(x, name)
        skips.append(x)
        x = F.max_pool2d(x, 2)
    x = block(x, 'conv_7')
    for name, skip in zip(ZF_BLOCKS[6:], skips[::-1]):
        x = block(torch.cat([F.interpolate(x, scale_factor=2, mode='nearest'), skip], 1), name)
    return F.conv2d(x, sd['conv_final.weight'], sd['conv_final.bias'])


def unet_forward(sd, x, abn=False, quant=None):
    """UNet / UNetABN in eval mode (lib/models/unet.py:79-107, unet_abn.py:80-107): double_conv = (conv3x3 -> BatchNorm ->
    ReLU) x 2, or (conv3x3 -> InPlaceABN [leaky 0.01]) x 2 with abn=True; MaxPool2d(2) down; nearest x2 upsampling +
    torch.cat([skip, upsampled]) up; Dropout2d inactive; 1x1 output conv."""
    q = quant if quant is not None else (lambda t: t)

    def norm_act(t, prefix):
        if abn:
            return inplace_abn_eval(t, sd, prefix)
        return F.relu(_bn_eval(t, sd, prefix))

    def double_conv(t, prefix):
        i2 = 2 if abn else 3
        t = norm_act(F.conv2d(q(t), q(sd[prefix + '.0.weight']), sd[prefix + '.0.bias'], padding=1), prefix + '.1')
        return norm_act(F.conv2d(q(t), q(sd[prefix + '.%d.weight' % i2]), sd[prefix + '.%d.bias' % i2], padding=1),
                        prefix + '.%d' % (i2 + 1))

    x1 = double_conv(x, 'inc.conv.conv')
    xs = [x1]
    for i in range(1, 5):
        xs.append(double_conv(F.max_pool2d(q(xs[-1]), 2), 'down%d.mpconv.1.conv' % i))
    t = xs[4]
    for i, skip in zip(range(1, 5), (xs[3], xs[2], xs[1], xs[0])):
        t = F.interpolate(q(t), scale_factor=2, mode='nearest')
        t = double_conv(torch.cat([q(skip), t], dim=1), 'up%d.conv.conv' % i)
    return F.conv2d(q(t), sd['outc.conv.weight'], sd['outc.conv.bias'])


def _bn_relu_conv(x, sd, prefix, quant=None, padding=1):
    """DenseLayer / TransitionDown front: BatchNorm2d(eval) -> ReLU -> conv (lib/models/tiramisu.py:9-19,47-59)."""
    y = F.relu(F.batch_norm(x, sd[prefix + '.norm.running_mean'], sd[prefix + '.norm.running_var'],
                            sd[prefix + '.norm.weight'], sd[prefix + '.norm.bias'], training=False, eps=1e-5))
    w, b = sd[prefix + '.conv.weight'], sd[prefix + '.conv.bias']
    if quant is not None:
        y, w = quant(y), quant(w)
    return F.conv2d(y, w, b, padding=padding)


def fcdensenet_forward(sd, x, down_blocks=(5, 5, 5, 5, 5), up_blocks=(5, 5, 5, 5, 5), bottleneck_layers=5, quant=None):
    """FCDenseNet.forward in eval mode (lib/models/tiramisu.py:168-184); returns the finalConv logits (the reference
    defines a LogSoftmax but never applies it).  `quant` models bf16 storage of every activation the device keeps."""
    q = quant if quant is not None else (lambda t: t)

    def dense_block(t, prefix, n_layers, upsample):
        new = []
        for k in range(n_layers):
            out = q(_bn_relu_conv(t, sd, '%s.layers.%d' % (prefix, k), quant))
            t = torch.cat([t, out], 1)
            new.append(out)
        return torch.cat(new, 1) if upsample else t

    w, b = sd['firstconv.weight'], sd['firstconv.bias']
    out = q(F.conv2d(q(x), q(w), b, padding=1))
    skips = []
    for i, n_layers in enumerate(down_blocks):
        out = dense_block(out, 'denseBlocksDown.%d' % i, n_layers, False)
        skips.append(out)
        out = F.max_pool2d(q(_bn_relu_conv(out, sd, 'transDownBlocks.%d' % i, quant, padding=0)), 2)
    out = dense_block(out, 'bottleneck.bottleneck', bottleneck_layers, True)
    for i, n_layers in enumerate(up_blocks):
        skip = skips.pop()
        w, b = sd['transUpBlocks.%d.convTrans.weight' % i], sd['transUpBlocks.%d.convTrans.bias' % i]
        up = q(F.conv_transpose2d(q(out), q(w), b, stride=2))
        oy, ox = (up.shape[2] - skip.shape[2]) // 2, (up.shape[3] - skip.shape[3]) // 2
        up = up[:, :, oy:oy + skip.shape[2], ox:ox + skip.shape[3]]
        out = dense_block(torch.cat([up, skip], 1), 'denseBlocksUp.%d' % i, n_layers, i < len(up_blocks) - 1)
    return F.conv2d(q(out), q(sd['finalConv.weight']), sd['finalConv.bias'])


RESNET34_BLOCKS = (3, 4, 6, 3)


def _bn_eval(x, sd, prefix):
    return F.batch_norm(x, sd[prefix + '.running_mean'], sd[prefix + '.running_var'], sd[prefix + '.weight'],
                        sd[prefix + '.bias'], training=False, eps=1e-5)


def inplace_abn_eval(x, sd, prefix, eps=1e-5, slope=0.01):
    """InPlaceABN in eval mode (lib/modules/abn/functions.py:62-100 with mapillary/inplace_abn's forward kernel, the
    un-vendored third-party backend): (x - mean) / sqrt(var + eps) * (|weight| + eps) + bias, then leaky_relu(slope).
    The |weight| + eps scale is that library's published forward (it keeps the transform invertible); no version is
    pinned by the reference and no test of the reference covers it -> parity unpinned."""
    mean, var = sd[prefix + '.running_mean'], sd[prefix + '.running_var']
    gamma = sd[prefix + '.weight'].abs() + eps
    y = (x - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + eps) * gamma.view(1, -1, 1, 1)
    return F.leaky_relu(y + sd[prefix + '.bias'].view(1, -1, 1, 1), slope)


def linknet34_forward(sd, x, quant=None):
    """LinkNet34.forward in eval mode (lib/models/linknet.py:65-90; encoder = torchvision resnet34 layers)."""
    q = quant if quant is not None else (lambda t: t)

    def conv(t, key, **kw):
        return F.conv2d(q(t), q(sd[key + '.weight']), sd.get(key + '.bias'), **kw)

    x = F.relu(_bn_eval(conv(x, 'firstconv', stride=2, padding=3), sd, 'firstbn'))
    x = F.max_pool2d(q(x), 3, 2, 1)
    feats = []
    for li, n_blocks in enumerate(RESNET34_BLOCKS):
        for b in range(n_blocks):
            pre = 'encoder%d.%d' % (li + 1, b)
            stride = 2 if (li > 0 and b == 0) else 1
            ident = x
            out = F.relu(_bn_eval(conv(x, pre + '.conv1', stride=stride, padding=1), sd, pre + '.bn1'))
            out = _bn_eval(conv(q(out), pre + '.conv2', padding=1), sd, pre + '.bn2')
            if pre + '.downsample.0.weight' in sd:
                ident = q(_bn_eval(conv(x, pre + '.downsample.0', stride=stride), sd, pre + '.downsample.1'))
            x = q(F.relu(out + ident))
        feats.append(x)
    e1, e2, e3, e4 = feats

    def decoder(t, name):
        t = q(inplace_abn_eval(conv(t, name + '.conv1'), sd, name + '.abn1'))
        t = F.conv_transpose2d(t, q(sd[name + '.deconv2.weight']), sd[name + '.deconv2.bias'], stride=2, padding=1)
        t = q(inplace_abn_eval(t, sd, name + '.abn2'))
        return inplace_abn_eval(conv(t, name + '.conv3'), sd, name + '.abn3')

    d4 = q(decoder(e4, 'decoder4') + e3)
    d3 = q(decoder(d4, 'decoder3') + e2)
    d2 = q(decoder(d3, 'decoder2') + e1)
    d1 = q(decoder(d2, 'decoder1'))
    f = F.leaky_relu(F.conv_transpose2d(d1, q(sd['finaldeconv1.weight']), sd['finaldeconv1.bias'], stride=2), 0.01)
    f = F.leaky_relu(conv(q(f), 'finalconv2'), 0.01)
    return conv(q(f), 'finalconv3', padding=1)



# ------------------------------------------------------------------------------------------------ InPlaceABN
def _abn_act_forward(y, activation, slope):
    if activation == 'leaky_relu':
        return F.leaky_relu(y, slope)
    if activation == 'elu':
        return F.elu(y)
    return y


def inplace_abn_forward(x, weight, bias, running_mean, running_var, training=True, momentum=0.1, eps=1e-5,
                        activation='leaky_relu', slope=0.01):
    """InPlaceABN.forward (lib/modules/abn/functions.py:62-100) with the arithmetic of the un-vendored backend
    (mapillary/inplace_abn, no pinned version -> parity unpinned): biased batch variance, running statistics updated with
    the unbiased one (:84-85), y = (x - mean) * rsqrt(var + eps) * (|weight| + eps) + bias, then the activation.
    Returns (z, var used, new running_mean, new running_var); nothing is modified in place."""
    dims = [0] + list(range(2, x.dim()))
    count = x.numel() // x.shape[1]
    shape = [1, -1] + [1] * (x.dim() - 2)
    if training:
        mean, var = x.mean(dim=dims), x.var(dim=dims, unbiased=False)
        running_mean = running_mean * (1 - momentum) + momentum * mean
        running_var = running_var * (1 - momentum) + momentum * var * count / (count - 1)
    else:
        mean, var = running_mean, running_var
    gamma = weight.abs() + eps if weight is not None else torch.ones_like(mean)
    beta = bias if bias is not None else torch.zeros_like(mean)
    y = (x - mean.view(shape)) * (torch.rsqrt(var + eps) * gamma).view(shape) + beta.view(shape)
    return _abn_act_forward(y, activation, slope), var, running_mean, running_var


def inplace_abn_backward(z, dz, var, weight, bias, training=True, eps=1e-5, activation='leaky_relu', slope=0.01):
    """InPlaceABN.backward (functions.py:102-122): undo the activation on (z, dz), edz / eydz reductions in training mode
    (zeros in eval mode, the reference's own shortcut at :110-112), then the backend's backward:
    dx = (dy - edz/count - xhat * eydz/count) * gamma * rsqrt(var + eps), dweight = eydz * sign(weight), dbias = edz."""
    dims = [0] + list(range(2, z.dim()))
    count = z.numel() // z.shape[1]
    shape = [1, -1] + [1] * (z.dim() - 2)
    if activation == 'leaky_relu':
        dy = torch.where(z < 0, dz * slope, dz)
        y = torch.where(z < 0, z / slope, z)
    elif activation == 'elu':
        dy = torch.where(z < 0, dz * (z + 1), dz)
        y = torch.where(z < 0, torch.log1p(z), z)
    else:
        dy, y = dz, z
    gamma = weight.abs() + eps if weight is not None else torch.ones_like(var)
    beta = bias if bias is not None else torch.zeros_like(var)
    xhat = (y - beta.view(shape)) / gamma.view(shape)
    if training:
        edz, eydz = dy.sum(dim=dims), (xhat * dy).sum(dim=dims)
    else:
        edz, eydz = torch.zeros_like(var), torch.zeros_like(var)
    dx = (dy - (edz / count).view(shape) - xhat * (eydz / count).view(shape)) * (gamma * torch.rsqrt(var + eps)).view(shape)
    if weight is None:
        return dx, None, None
    return dx, torch.where(weight > 0, eydz, -eydz), edz



def linknet34_forward_train(sd, x, quant=None, linear=False, keep=None, drop_p=0.5):
    """LinkNet34.forward in train() mode (lib/models/linknet.py:65-90): every BatchNorm2d / InPlaceABN normalises with
    batch statistics and updates its running statistics (momentum 0.1, unbiased variance).  `keep` = the Dropout2d keep
    mask [N, 64] of finaldrop1 (:57,83: whole channels of decoder1's output are zeroed, the rest scaled by 1 / (1 - p));
    None = dropout inactive.
    Returns (logits, {buffer name: updated value}); `sd` is not modified.  linear=True drops every ReLU / leaky-ReLU (a
    test mode: without activation gates the gradients are not chaotic under bf16 rounding, so the whole backward graph
    can be compared tightly)."""
    relu = (lambda t: t) if linear else F.relu
    leaky = (lambda t, s: t) if linear else F.leaky_relu
    q = quant if quant is not None else (lambda t: t)
    new = {}

    def conv(t, key, **kw):
        return F.conv2d(q(t), q(sd[key + '.weight']), sd.get(key + '.bias'), **kw)

    def bn(t, prefix):
        rm, rv = sd[prefix + '.running_mean'].clone(), sd[prefix + '.running_var'].clone()
        y = F.batch_norm(t, rm, rv, sd[prefix + '.weight'], sd[prefix + '.bias'], training=True, momentum=0.1, eps=1e-5)
        new[prefix + '.running_mean'], new[prefix + '.running_var'] = rm, rv
        return y

    def abn(t, prefix):
        z, _, rm, rv = inplace_abn_forward(t, sd[prefix + '.weight'], sd[prefix + '.bias'], sd[prefix + '.running_mean'],
                                           sd[prefix + '.running_var'], True, 0.1, 1e-5, 'none' if linear else 'leaky_relu', 0.01)
        new[prefix + '.running_mean'], new[prefix + '.running_var'] = rm, rv
        return z

    x = relu(bn(conv(x, 'firstconv', stride=2, padding=3), 'firstbn'))
    x = F.max_pool2d(q(x), 3, 2, 1)
    feats = []
    for li, n_blocks in enumerate(RESNET34_BLOCKS):
        for b in range(n_blocks):
            pre = 'encoder%d.%d' % (li + 1, b)
            stride = 2 if (li > 0 and b == 0) else 1
            ident = x
            out = q(relu(bn(q(conv(x, pre + '.conv1', stride=stride, padding=1)), pre + '.bn1')))
            out = bn(q(conv(out, pre + '.conv2', padding=1)), pre + '.bn2')
            if pre + '.downsample.0.weight' in sd:
                ident = q(bn(q(conv(x, pre + '.downsample.0', stride=stride)), pre + '.downsample.1'))
            x = q(relu(out + ident))
        feats.append(x)
    e1, e2, e3, e4 = feats

    def decoder(t, name):
        t = q(abn(q(conv(t, name + '.conv1')), name + '.abn1'))
        t = F.conv_transpose2d(t, q(sd[name + '.deconv2.weight']), sd[name + '.deconv2.bias'], stride=2, padding=1)
        t = q(abn(q(t), name + '.abn2'))
        return abn(q(conv(t, name + '.conv3')), name + '.abn3')

    d4 = q(decoder(e4, 'decoder4') + e3)
    d3 = q(decoder(d4, 'decoder3') + e2)
    d2 = q(decoder(d3, 'decoder2') + e1)
    d1 = q(decoder(d2, 'decoder1'))
    if keep is not None:
        d1 = q(d1 * (keep.to(d1.dtype) / (1.0 - drop_p)).view(d1.shape[0], 64, 1, 1))
    f = leaky(F.conv_transpose2d(d1, q(sd['finaldeconv1.weight']), sd['finaldeconv1.bias'], stride=2), 0.01)
    f = leaky(conv(q(f), 'finalconv2'), 0.01)
    return conv(q(f), 'finalconv3', padding=1), new


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


# ------------------------------------------------------------------------------------------- loss / metrics
def bce_with_sigmoid(logits, targets):
    """lib/losses.py:46-53: BCE-with-logits applied to logsigmoid(logits) (the reference's double squash), mean."""
    z = F.logsigmoid(logits)
    return F.binary_cross_entropy_with_logits(z, targets.float(), reduction='mean')


def smooth_jaccard(logits, targets, smooth=100):
    p = torch.sigmoid(logits)
    t = targets.float()
    inter = torch.sum(p * t)
    union = torch.sum(p) + torch.sum(t)
    return 1 - (inter + smooth) / (union - inter + smooth)


def bce_elements(logits, targets):
    """lib/losses.py:46-53 with reduce=False: the per-element loss tensor."""
    return F.binary_cross_entropy_with_logits(F.logsigmoid(logits), targets.float(), reduction='none')


def jaccard_loss(logits, targets):
    """lib/losses.py:18-28."""
    p = torch.sigmoid(logits)
    inter = torch.sum(p * targets)
    union = torch.sum(p) + torch.sum(targets)
    return 1 - inter / (union - inter + 1e-7)


def focal_loss_binary(logits, targets, gamma=2, size_average=True):
    """lib/losses.py:78-101: logpt = -BCE-with-logits(logsigmoid(x), t); loss = -(1 - exp(logpt))^gamma * logpt."""
    logpt = -bce_elements(logits, targets)
    pt = torch.exp(logpt)
    loss = -((1 - pt).pow(gamma)) * logpt
    return loss.mean() if size_average else loss.sum()


def bce_jaccard(logits, targets, bce_weight=1, jaccard_weight=0.5):
    return (bce_with_sigmoid(logits, targets) * bce_weight + smooth_jaccard(logits, targets) * jaccard_weight) / (
        bce_weight + jaccard_weight)


def jaccard_score(logits, targets):
    p = torch.sigmoid(logits)
    t = targets.float()
    inter = (p * t).sum()
    union = p.sum() + t.sum()
    return inter / (union - inter + 1e-7)


def pixel_accuracy(logits, targets):
    pred = torch.sigmoid(logits) > 0.5
    n_true = torch.eq(pred, targets.byte()).sum()
    if n_true == 0:
        return n_true
    return n_true.float() / targets.numel()


def confusion_counts(probs, targets, thr=0.5):
    """int64 [tp, fp, fn, tn] at probs > thr."""
    pred = (probs > thr).reshape(-1)
    truth = (targets != 0).reshape(-1)
    return torch.stack([(pred & truth).sum(), (pred & ~truth).sum(), (~pred & truth).sum(), (~pred & ~truth).sum()])


def pr_curve_counts(logits, targets, n_thresholds=127):
    """lib/train_utils.py:92-125 -> uint64 arrays tp, tn, fp, fn per threshold arange(0, 1, 1/n) (float32)."""
    thr = np.arange(0., 1., 1. / n_thresholds, dtype=np.float32)
    p = torch.sigmoid(logits).numpy().reshape(-1)
    t = targets.numpy().astype(np.int32).reshape(-1)
    tp = np.zeros(len(thr), np.uint64); tn = np.zeros(len(thr), np.uint64)
    fp = np.zeros(len(thr), np.uint64); fn = np.zeros(len(thr), np.uint64)
    for i, v in enumerate(thr):
        conf = np.bincount((p > v).astype(np.int32) + 2 * t, minlength=4).reshape(2, 2)   # [truth][pred]
        tp[i], tn[i], fp[i], fn[i] = conf[1, 1], conf[0, 0], conf[0, 1], conf[1, 0]
    return tp, tn, fp, fn

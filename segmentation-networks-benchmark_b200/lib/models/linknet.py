"""LinkNet34 (reference lib/models/linknet.py:33-90): ResNet-34 encoder + LinkNet decoders with InPlaceABN.

Same constructor and state_dict keys as the reference (`firstconv`, `firstbn`, `encoder1..4` = torchvision resnet34
layers, `decoder1..4.{conv1,abn1,deconv2,abn2,conv3,abn3}`, `finaldeconv1`, `finalconv2`, `finalconv3`).  The encoder
is rebuilt here with torchvision's module names so no torchvision import is needed; `pretrained=True` cannot download
weights in this environment and, like every other model here, starts from random initialisation.

Forward runs on the native engine in eval mode (BatchNorm / InPlaceABN folded into the convolutions) and in train mode
(batch statistics, running statistics updated in place, Dropout2d as a per-(image, channel) keep mask).  In train mode
with grad enabled the forward is one autograd node whose backward runs the whole backward pass of BASELINE configs[1] on
the tensor cores (snb_b200.train_engine.LinkNet34TrainPlan: tcgen05 input and weight gradients, BatchNorm / InPlaceABN
backward) and hands every parameter its gradient; `train_step` fuses forward, loss and backward.  `InPlaceABN` holds the
parameters of
mapillary's in-place activated batch norm (lib/modules/abn/bn.py:47-103); in eval mode it is
leaky_relu((x - mean) / sqrt(var + eps) * (|weight| + eps) + bias, 0.01) -- the `|weight| + eps` scale is that
library's forward; the library is not vendored in the reference and has no pinned version (parity unpinned).
"""
import torch
from torch import nn

from ... import _native as N
from ...engine import LinkNet34Plan
from ...train_engine import LinkNet34TrainPlan
from ..modules.abn import InPlaceABN


class BasicBlock(nn.Module):
    """torchvision.models.resnet.BasicBlock parameter holder (conv1, bn1, conv2, bn2, optional downsample)."""

    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = None
        if stride != 1 or inplanes != planes:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes, 1, stride=stride, bias=False), nn.BatchNorm2d(planes))
        self.stride = stride


def _resnet_layer(inplanes, planes, blocks, stride):
    return nn.Sequential(*[BasicBlock(inplanes if i == 0 else planes, planes, stride if i == 0 else 1) for i in range(blocks)])


class DecoderBlockLinkNet(nn.Module):
    def __init__(self, in_channels, n_filters):
        super().__init__()
        q = in_channels // 4
        self.conv1, self.abn1 = nn.Conv2d(in_channels, q, 1), InPlaceABN(q)
        self.deconv2 = nn.ConvTranspose2d(q, q, kernel_size=4, stride=2, padding=1, output_padding=0)
        self.abn2 = InPlaceABN(q)
        self.conv3, self.abn3 = nn.Conv2d(q, n_filters, 1), InPlaceABN(n_filters)


def _bn(m, abn=False):
    return (m.weight, m.bias, m.running_mean, m.running_var, m.eps, abn)


class _TrainStep(torch.autograd.Function):
    """train() forward + backward of LinkNet34 as one autograd node: forward runs LinkNet34TrainPlan, backward walks its
    tape (BatchNorm / ABN backward, generic dgrad / wgrad, max-pool backward) and hands every parameter its gradient."""

    @staticmethod
    def forward(ctx, x, model, *params):
        plan = model.plan_train(x.shape[0], x.shape[2], x.shape[3])
        plan.load_nchw(x.detach().float())
        out = plan.run()
        # the plan's slabs ARE the saved activations: a later forward of the same shape overwrites them
        ctx.plan, ctx.generation, ctx.model_params = plan, plan.generation, list(model.parameters())
        return out.unsqueeze(1).clone()

    @staticmethod
    def backward(ctx, dout):
        plan = ctx.plan
        if plan.generation != ctx.generation:
            raise RuntimeError("LinkNet34: another train-mode forward of the same input shape ran before this backward; its "
                               "activations replaced the ones this graph saved (run forward and backward in pairs)")
        with torch.cuda.device(dout.device):
            plan.backward(dout.contiguous())
            pg = plan.grads_in(plan.grad_arena.clone())       # autograd may keep (or accumulate into) what it is handed
        return (None, None) + tuple(pg.get(p) for p in ctx.model_params)


class LinkNet34(nn.Module):
    def __init__(self, num_classes=1, num_channels=3, pretrained=True):
        super().__init__()
        assert num_channels == 3
        self.num_classes = num_classes
        filters = [64, 128, 256, 512]
        self.firstconv = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.firstbn = nn.BatchNorm2d(64)
        self.firstrelu = nn.ReLU(inplace=True)
        self.firstmaxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        for i, (blocks, stride) in enumerate([(3, 1), (4, 2), (6, 2), (3, 2)]):
            setattr(self, 'encoder%d' % (i + 1), _resnet_layer(filters[max(i - 1, 0)], filters[i], blocks, stride))
        for i in range(4, 0, -1):
            setattr(self, 'decoder%d' % i, DecoderBlockLinkNet(filters[i - 1], filters[max(i - 2, 0)]))
        self.finaldrop1 = nn.Dropout2d(p=0.5)
        self.finaldeconv1 = nn.ConvTranspose2d(filters[0], 32, 3, stride=2)
        self.finalrelu1 = nn.LeakyReLU(inplace=True)
        self.finalconv2 = nn.Conv2d(32, 32, 3)
        self.finalrelu2 = nn.LeakyReLU(inplace=True)
        self.finalconv3 = nn.Conv2d(32, num_classes, 2, padding=1)

    def _spec(self):
        def block(b):
            down = None if b.downsample is None else (b.downsample[0].weight, _bn(b.downsample[1]))
            return dict(conv1=b.conv1.weight, bn1=_bn(b.bn1), conv2=b.conv2.weight, bn2=_bn(b.bn2), down=down)

        def dec(d):
            return dict(conv1=(d.conv1.weight, d.conv1.bias), abn1=_bn(d.abn1, True),
                        deconv2=(d.deconv2.weight, d.deconv2.bias), abn2=_bn(d.abn2, True),
                        conv3=(d.conv3.weight, d.conv3.bias), abn3=_bn(d.abn3, True))

        return dict(stem=(self.firstconv.weight, _bn(self.firstbn)),
                    encoders=[[block(b) for b in getattr(self, 'encoder%d' % i)] for i in range(1, 5)],
                    decoders=[dec(getattr(self, 'decoder%d' % i)) for i in range(1, 5)],
                    final1=(self.finaldeconv1.weight, self.finaldeconv1.bias),
                    final2=(self.finalconv2.weight, self.finalconv2.bias),
                    final3=(self.finalconv3.weight, self.finalconv3.bias))

    def _stamp(self):
        tensors = list(self.parameters()) + list(self.buffers())
        return tuple((t.data_ptr(), t._version) for t in tensors)

    def plan(self, n, h, w, sigmoid=False):
        cache = self.__dict__.setdefault('_plans', {})
        stamp = self._stamp()
        if self.__dict__.get('_plan_stamp') != stamp:
            cache.clear()
            self.__dict__['_plan_stamp'] = stamp
        key = (n, h, w, bool(sigmoid))
        if key not in cache:
            dev = self.finalconv3.weight.device
            if dev.type != 'cuda':
                raise RuntimeError("LinkNet34 runs on CUDA devices only (no CPU fallback); call .cuda()")
            with torch.no_grad():
                cache[key] = LinkNet34Plan(self._spec(), n, h, w, dev, sigmoid)
        return cache[key]

    def plan_train(self, n, h, w):
        """Training-mode plan (batch statistics).  After an optimiser step (same parameter storages, new versions) the
        cached plans re-pack their weights in place (`LinkNet34TrainPlan.refresh`); a plan is only rebuilt when a parameter
        was replaced by a different tensor."""
        cache = self.__dict__.setdefault('_train_plans', {})
        ptrs = tuple(p.data_ptr() for p in self.parameters())
        versions = tuple(p._version for p in self.parameters()) + tuple(
            b._version for k, b in self.named_buffers() if not k.endswith('num_batches_tracked'))
        if self.__dict__.get('_train_ptrs') != ptrs:
            cache.clear()
            self.__dict__['_train_ptrs'] = ptrs
            self.__dict__['_train_versions'] = {}
        key = (n, h, w)
        if key not in cache:
            dev = self.finalconv3.weight.device
            if dev.type != 'cuda':
                raise RuntimeError("LinkNet34 runs on CUDA devices only (no CPU fallback); call .cuda()")
            with torch.no_grad():
                cache[key] = LinkNet34TrainPlan(self, n, h, w, dev, linear=bool(self.__dict__.get('_test_linear', False)))
            self.__dict__['_train_versions'][key] = versions
        elif self.__dict__['_train_versions'].get(key) != versions:
            cache[key].refresh()
            self.__dict__['_train_versions'][key] = versions
        return cache[key]

    def train_step(self, x, targets, criterion, loss_scale=None):
        """One training step without the optimiser (torch_train.py:183-189: `outputs = model(x); loss = criterion(outputs, y);
        (batch_size * loss).backward()`), fused: train-mode forward (one CUDA-graph replay), loss reduction and its gradient
        (two launches, no autograd, no host synchronisation), backward (one CUDA-graph replay).  Every parameter's .grad is
        set to its slice of the plan's gradient arena (valid until the next step of this shape).  `criterion` is one of
        the fused losses of snb_b200.lib.losses; loss_scale defaults to the batch size as in the reference's loop.
        Returns (loss, logits [N, 1, H, W]) -- device tensors, nothing is synchronised."""
        from ..losses import fused_loss_and_grad

        N.require_cuda()
        if not self.training:
            raise RuntimeError("train_step needs train() mode")
        if not x.is_cuda or x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("expected a CUDA input of shape [N, 3, H, W]")
        with torch.cuda.device(x.device), torch.no_grad():
            self.__dict__.pop('_plan_stamp', None)
            plan = self.plan_train(x.shape[0], x.shape[2], x.shape[3])
            plan.load_nchw(x.detach().float())
            logits = plan.run().unsqueeze(1)
            loss, _, _ = fused_loss_and_grad(criterion, logits, targets, plan.dlogits,
                                             float(x.shape[0]) if loss_scale is None else loss_scale)
            grads = plan.backward(None)
            for p in self.parameters():
                p.grad = grads[p] if p.requires_grad else None
        return loss, logits

    def forward(self, x):
        """eval(): folded BatchNorm / InPlaceABN (running statistics).  train(): batch statistics, running statistics
        updated in place; with grad enabled the logits are an autograd node whose backward fills every parameter's
        .grad (LinkNet34TrainPlan.backward).  The gradient with respect to the input image is not computed."""
        N.require_cuda()
        if not x.is_cuda:
            raise RuntimeError("input must be a CUDA tensor (no CPU fallback)")
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("expected input of shape [N, 3, H, W]")
        with torch.cuda.device(x.device):
            if self.training:
                self.__dict__.pop('_plan_stamp', None)        # the eval plans fold the running statistics: rebuild them
                if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
                    return _TrainStep.apply(x, self, *self.parameters())
                p = self.plan_train(x.shape[0], x.shape[2], x.shape[3])
            else:
                p = self.plan(x.shape[0], x.shape[2], x.shape[3], sigmoid=False)
            p.load_nchw(x.detach().float())
            out = p.run()
        return out.unsqueeze(1).clone()

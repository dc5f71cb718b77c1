"""UNet16 = TernausNet-VGG16, the "AlbuNet" of BASELINE.json (reference lib/models/unet16.py:52-131).

Same constructor, attribute names and state_dict keys (every encoder conv appears under `encoder.N.*` and
`convK.M.*`, 76 entries / 50 tensors); the forward pass runs on the native sm_100a engine.
"""
from torch import nn

from ._vgg_unet import ConvRelu, DecoderBlock, VGGUNetBase, vgg_features

_VGG16 = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M']


class UNet16(VGGUNetBase):
    def __init__(self, num_classes=1, num_filters=32, pretrained=False):
        super().__init__()
        self.num_classes = num_classes
        self.pool = nn.MaxPool2d(2, 2)
        # pretrained == 'vgg' would download ImageNet weights in the reference; there is no network here and
        # get_model passes True, which the reference ignores (torch_train.py:113, unet16.py:66) -> random init
        self.encoder = vgg_features(_VGG16)
        self.relu = nn.ReLU(inplace=True)
        e = self.encoder
        self.conv1 = nn.Sequential(e[0], self.relu, e[2], self.relu)
        self.conv2 = nn.Sequential(e[5], self.relu, e[7], self.relu)
        self.conv3 = nn.Sequential(e[10], self.relu, e[12], self.relu, e[14], self.relu)
        self.conv4 = nn.Sequential(e[17], self.relu, e[19], self.relu, e[21], self.relu)
        self.conv5 = nn.Sequential(e[24], self.relu, e[26], self.relu, e[28], self.relu)

        nf = num_filters
        # (name, channels in, middle, out): decoder inputs are [previous decoder output | encoder skip]
        for name, c_in, mid, out in [('center', 512, nf * 16, nf * 8), ('dec5', 512 + nf * 8, nf * 16, nf * 8),
                                     ('dec4', 512 + nf * 8, nf * 16, nf * 8), ('dec3', 256 + nf * 8, nf * 8, nf * 2),
                                     ('dec2', 128 + nf * 2, nf * 4, nf)]:
            setattr(self, name, DecoderBlock(c_in, mid, out))
        self.dec1 = ConvRelu(64 + nf, nf)
        self.final = nn.Conv2d(nf, num_classes, kernel_size=1)

    def _stages(self):
        e = self.encoder
        return [[e[0], e[2]], [e[5], e[7]], [e[10], e[12], e[14]], [e[17], e[19], e[21]], [e[24], e[26], e[28]]]

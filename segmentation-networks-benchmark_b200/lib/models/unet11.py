"""UNet11 = TernausNet-VGG11 (reference lib/models/unet11.py:51-122) on the native sm_100a engine."""
from torch import nn

from ._vgg_unet import ConvRelu, DecoderBlock, VGGUNetBase, vgg_features

_VGG11 = [64, 'M', 128, 'M', 256, 256, 'M', 512, 512, 'M', 512, 512, 'M']


class UNet11(VGGUNetBase):
    def __init__(self, num_classes=1, num_filters=32, pretrained=False):
        super().__init__()
        self.pool = nn.MaxPool2d(2, 2)
        self.num_classes = num_classes
        self.encoder = vgg_features(_VGG11)
        self.relu = nn.ReLU(inplace=True)
        e = self.encoder
        self.conv1 = nn.Sequential(e[0], self.relu)
        self.conv2 = nn.Sequential(e[3], self.relu)
        self.conv3 = nn.Sequential(e[6], self.relu, e[8], self.relu)
        self.conv4 = nn.Sequential(e[11], self.relu, e[13], self.relu)
        self.conv5 = nn.Sequential(e[16], self.relu, e[18], self.relu)

        nf = num_filters
        # (name, channels in, middle, out): decoder inputs are [previous decoder output | encoder skip]
        for name, c_in, mid, out in [('center', 256 + nf * 8, nf * 16, nf * 8), ('dec5', 512 + nf * 8, nf * 16, nf * 8),
                                     ('dec4', 512 + nf * 8, nf * 16, nf * 4), ('dec3', 256 + nf * 4, nf * 8, nf * 2),
                                     ('dec2', 128 + nf * 2, nf * 4, nf)]:
            setattr(self, name, DecoderBlock(c_in, mid, out, is_deconv=True))
        self.dec1 = ConvRelu(64 + nf, nf)
        self.final = nn.Conv2d(nf, num_classes, kernel_size=1)

    def _stages(self):
        e = self.encoder
        return [[e[0]], [e[3]], [e[6], e[8]], [e[11], e[13]], [e[16], e[18]]]

"""nn.Module mirrors of the reference's lib/models for the hot path (same class names and state_dict keys)."""
from .unet11 import UNet11  # noqa: F401
from .unet16 import UNet16  # noqa: F401
from .zf_unet import ZF_UNET  # noqa: F401
from .tiramisu import FCDenseNet, FCDenseNet57, FCDenseNet67, FCDenseNet103  # noqa: F401
from .linknet import LinkNet34  # noqa: F401
from .unet import UNet, UNetABN  # noqa: F401

from .blocks import ConvBlock, ResNetBlock  # noqa: F401
from .discriminator import PatchGANDiscriminator  # noqa: F401
from .generator import ResnetGenerator  # noqa: F401
from .loss import HULoss, WassersteinLoss, ZNCCLoss  # noqa: F401

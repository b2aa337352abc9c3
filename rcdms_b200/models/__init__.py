from .unet import UNet3DConditionModel, UNet3DConditionOutput  # noqa: F401
from .myprior_transformer import MyPriorTransformer, PriorTransformerOutput  # noqa: F401
from .vae import AutoencoderKL  # noqa: F401

from .unet import UNet3DConditionModel, UNet3DConditionOutput  # noqa: F401

from rcdms_b200.models.unet import UNet3DConditionModel, UNet3DConditionOutput  # noqa: F401

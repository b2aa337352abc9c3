"""Import-compatibility shim: the reference's callers do ``from src.models.unet import UNet3DConditionModel`` and
``from src.pipelines.RCDMs_pipeline import RCDMsPipeline`` (stage2_batchtest_rcdms_model.py:8,30).  These modules
re-export the B200-native drop-ins so such callers run unchanged."""

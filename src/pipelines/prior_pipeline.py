"""Import path of the reference (``src/pipelines/prior_pipeline.py``): the B200-native drop-in."""
from rcdms_b200.pipelines.prior_pipeline import KandinskyPriorPipelineOutput, Seq_Inpaint_Prior_Pipeline  # noqa: F401

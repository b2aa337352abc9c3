from rcdms_b200.pipelines.RCDMs_pipeline import RCDMsPipeline, RCDMsPipelineOutput, local_feature  # noqa: F401

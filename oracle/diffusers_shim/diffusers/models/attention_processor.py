"""Names imported by myprior_transformer.py (stage 1; not on the stage-2 path)."""
ADDED_KV_ATTENTION_PROCESSORS = ()
CROSS_ATTENTION_PROCESSORS = ()
AttentionProcessor = object
AttnAddedKVProcessor = object
AttnProcessor = object

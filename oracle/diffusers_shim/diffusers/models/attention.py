from oracle.diffusers_restated import FeedForward, AdaLayerNorm, GEGLU, GELU  # noqa: F401

from oracle.diffusers_restated import Timesteps, TimestepEmbedding  # noqa: F401

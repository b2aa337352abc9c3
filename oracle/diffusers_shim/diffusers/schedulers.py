from oracle.diffusers_restated import DDIMSchedulerRef as DDIMScheduler  # noqa: F401

"""TEST INFRASTRUCTURE ONLY — minimal stand-in for ``diffusers`` (absent from this image).

Exists so that the reference's own ``/root/reference/src/models/*.py`` import and run
unmodified inside the build container when ``oracle/make_golden.py`` generates golden
vectors.  All arithmetic lives in ``oracle/diffusers_restated.py``.  Never on a product path.
"""
from .models.modeling_utils import ModelMixin  # noqa: F401
from .configuration_utils import ConfigMixin, register_to_config  # noqa: F401


class DiffusionPipeline:  # placeholder: the reference pipeline module is not imported by the oracle
    pass

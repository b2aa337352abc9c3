"""Import path of the reference (``src/models/myprior_transformer.py``): the B200-native drop-in."""
from rcdms_b200.models.myprior_transformer import MyPriorTransformer, PriorTransformerOutput  # noqa: F401

"""Register this package's BLIP2_MR / Blip2T5 in the reference's own LAVIS registry.
Importing this module requires the reference package (`lavis`) to be importable."""
from lavis.common.registry import registry as lavis_registry  # noqa: E402

from .blip2_mr import BLIP2_MR
from .blip2_t5 import Blip2T5

lavis_registry.mapping["model_name_mapping"]["blip2_mr"] = BLIP2_MR
lavis_registry.mapping["model_name_mapping"]["blip2_t5"] = Blip2T5
__all__ = ["BLIP2_MR", "Blip2T5"]

"""Attribution patching (reference features/patching): which SAE latents matter for a logit difference."""
from .attribution import Attribution, attribution_for_feature
from .utils import get_logit_diff, get_model_backward_cache_with_sae, get_model_forward_cache_with_sae

__all__ = ["Attribution", "attribution_for_feature", "get_logit_diff", "get_model_forward_cache_with_sae",
           "get_model_backward_cache_with_sae"]

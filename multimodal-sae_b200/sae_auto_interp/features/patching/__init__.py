"""Attribution patching (reference features/patching): which SAE latents matter for a logit difference."""
from . import attribution as _attribution, utils as _utils

Attribution, attribution_for_feature = _attribution.Attribution, _attribution.attribution_for_feature
get_logit_diff = _utils.get_logit_diff
get_model_forward_cache_with_sae = _utils.get_model_forward_cache_with_sae
get_model_backward_cache_with_sae = _utils.get_model_backward_cache_with_sae
__all__ = ["Attribution", "attribution_for_feature", "get_logit_diff", "get_model_forward_cache_with_sae",
           "get_model_backward_cache_with_sae"]

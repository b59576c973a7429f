"""`sae_auto_interp.features`: cache writer, cache reader, example constructors, samplers, steering (hot-path mirror)."""
from . import cache as _cache, constructors as _ctor, features as _types, loader as _loader, samplers as _samplers
from . import steering as _steering
from .patching import Attribution

FeatureCache, FeatureImageCache = _cache.FeatureCache, _cache.FeatureImageCache
FeatureDataset = _loader.FeatureDataset
Feature, FeatureRecord, Example = _types.Feature, _types.FeatureRecord, _types.Example
(pool_max_activation_windows, pool_max_activations_windows_image, random_activation_windows,
 random_activations_image, default_constructor, top_windows_all_features, top_images_all_features) = (
    _ctor.pool_max_activation_windows, _ctor.pool_max_activations_windows_image, _ctor.random_activation_windows,
    _ctor.random_activations_image, _ctor.default_constructor, _ctor.top_windows_all_features,
    _ctor.top_images_all_features)
sample, sample_with_explanation = _samplers.sample, _samplers.sample_with_explanation
SteeringController = _steering.SteeringController

# same public names as the reference package (its stats helpers are outside the accelerated path)
__all__ = ("FeatureCache FeatureImageCache FeatureDataset Feature FeatureRecord Example "
           "pool_max_activation_windows pool_max_activations_windows_image random_activation_windows "
           "random_activations_image default_constructor top_windows_all_features top_images_all_features sample sample_with_explanation "
           "SteeringController Attribution").split()

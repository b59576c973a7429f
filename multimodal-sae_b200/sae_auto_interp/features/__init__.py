from .cache import FeatureCache, FeatureImageCache
from .constructors import (
    default_constructor,
    pool_max_activation_windows,
    pool_max_activations_windows_image,
    random_activation_windows,
    random_activations_image,
    top_windows_all_features,
)
from .features import Example, Feature, FeatureRecord
from .loader import FeatureDataset
from .samplers import sample, sample_with_explanation
from .steering import SteeringController

__all__ = [
    "FeatureCache",
    "FeatureImageCache",
    "FeatureDataset",
    "Feature",
    "FeatureRecord",
    "Example",
    "pool_max_activation_windows",
    "pool_max_activations_windows_image",
    "random_activation_windows",
    "random_activations_image",
    "default_constructor",
    "top_windows_all_features",
    "sample",
    "sample_with_explanation",
    "SteeringController",
]

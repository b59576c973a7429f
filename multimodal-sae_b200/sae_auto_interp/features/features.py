"""Record types handed to explainers / scorers (reference features/features.py).  Only the data carriers are kept;
image mask compositing (PIL) is host-side presentation code outside the accelerated path."""
from __future__ import annotations

import json
from dataclasses import dataclass
from typing import Any, List, Optional


@dataclass
class Example:
    tokens: Any  # Tensor[ctx]
    activations: Any  # Tensor[ctx]

    def __hash__(self) -> int:
        return hash(tuple(self.tokens.tolist()))

    def __eq__(self, other: "Example") -> bool:
        return self.tokens.tolist() == other.tokens.tolist()

    @property
    def max_activation(self):
        return max(self.activations)


@dataclass
class ImageExample(Example):
    image: Any = None
    activation_image: Any = None
    mask: Any = None


def prepare_examples(tokens, activations) -> List[Example]:
    return [Example(tokens=t, activations=a) for t, a in zip(tokens, activations)]


@dataclass
class Feature:
    module_name: str
    feature_index: int

    def __repr__(self) -> str:
        return f"{self.module_name}_feature{self.feature_index}"


class FeatureRecord:
    def __init__(self, feature: Feature):
        self.feature = feature
        self.train: Optional[list] = None
        self.explanation: Optional[str] = None
        self.examples: Optional[list] = None

    @property
    def max_activation(self):
        return self.examples[0].max_activation

    def save(self, directory: str, save_examples: bool = False) -> None:
        payload = dict(self.__dict__)
        if not save_examples:
            for key in ("examples", "train", "test"):
                payload.pop(key, None)
        payload.pop("feature", None)
        with open(f"{directory}/{self.feature}.json", "w") as fh:
            json.dump(payload, fh, default=str)

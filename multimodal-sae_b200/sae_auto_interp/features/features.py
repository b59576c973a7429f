"""Record types handed to explainers / scorers (reference features/features.py): the data carriers, and the image
examples (activation mask upsampled onto the resized image) that `explain_images` shows to the explainer."""
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


def upsample_mask(mask, image_size, value: int = 224):
    """[p, p] activations -> PIL "L" mask of `image_size`: `value` where the patch is inactive (< 1e-5), 0 where it
    fired, bilinearly resized (reference features.py:134-141)."""
    import numpy as np
    from PIL import Image

    grey = ((mask < 1e-5).to(dtype=int).numpy() * value).astype(np.uint8)
    return Image.fromarray(grey, mode="L").resize(image_size, Image.BILINEAR)


def prepare_image_examples(tokens, activations, images, processor=None) -> List[ImageExample]:
    """One `ImageExample` per (token row, activation row, PIL image): the first `num_image_tokens` activations are the
    base image patches (24 x 24 at 576 tokens -> 336 px, else 27 x 27 -> 384 px); inactive regions of the resized
    image are blacked out through the upsampled mask (reference features.py:49-92)."""
    from PIL import Image

    n_base = getattr(processor, "num_image_tokens", 576)
    patches = 24 if n_base == 576 else 27
    side = 336 if patches == 24 else 384
    black = Image.new("L", (side, side), 0).convert("RGB")
    out = []
    for toks, acts, img in zip(tokens, activations, images):
        mask = upsample_mask(acts[:n_base].view(patches, patches), (side, side))
        shown = Image.composite(black, img.resize((side, side)), mask).convert("RGB")
        out.append(ImageExample(tokens=toks, activations=acts, image=img, activation_image=shown, mask=mask))
    return out


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

"""Stand-ins for optional third-party helpers the reference imports (simple_parsing, natsort)."""
from __future__ import annotations

import dataclasses
import re

try:  # the reference's config base class; optional here
    from simple_parsing import Serializable, field, list_field  # type: ignore
except Exception:  # pragma: no cover - exercised where simple_parsing is absent

    class Serializable:
        """Just enough of simple_parsing.Serializable for cfg.json round trips."""

        def to_dict(self) -> dict:
            return dataclasses.asdict(self)

        @classmethod
        def from_dict(cls, d: dict):
            names = {f.name for f in dataclasses.fields(cls)}
            return cls(**{k: v for k, v in d.items() if k in names})

    def field(default=dataclasses.MISSING, default_factory=dataclasses.MISSING, **_ignored):
        if default_factory is not dataclasses.MISSING:
            return dataclasses.field(default_factory=default_factory)
        if default is not dataclasses.MISSING:
            return dataclasses.field(default=default)
        return dataclasses.field()

    def list_field(*values, **_ignored):
        return dataclasses.field(default_factory=lambda: list(values))


try:
    from natsort import natsorted  # type: ignore
except Exception:  # pragma: no cover

    def natsorted(seq, key=None):
        def nat_key(item):
            text = str(key(item) if key is not None else item)
            return [int(tok) if tok.isdigit() else tok for tok in re.split(r"(\d+)", text)]

        return sorted(seq, key=nat_key)


def table_dataclass(name: str, doc: str, module: str, *rows, positional=(), post_init=None):
    """A `Serializable` dataclass from a table of (field name, type, default) rows -- `dataclasses.MISSING` marks a
    required field, a list default becomes a `list_field`, `positional` names take no flag on the command line."""
    fields = []
    for fname, ftype, default in rows:
        if default is dataclasses.MISSING:
            fields.append((fname, ftype))
        elif fname in positional:
            fields.append((fname, ftype, field(default=default, positional=True)))
        elif isinstance(default, list):
            fields.append((fname, ftype, list_field(*default)))
        else:
            fields.append((fname, ftype, field(default=default)))
    namespace = {} if post_init is None else {"__post_init__": post_init}
    cls = dataclasses.make_dataclass(name, fields, bases=(Serializable,), namespace=namespace)
    cls.__doc__, cls.__module__ = doc, module
    return cls

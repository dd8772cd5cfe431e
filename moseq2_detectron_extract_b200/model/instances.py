"""Minimal stand-in for detectron2.structures.Instances / Boxes with the operations the extract steps use
(len, x[i], x[bool tensor], x[list], .to(), .image_size, Instances.cat, fields).
Tensors may live on the GPU; nothing here forces a host copy."""
from __future__ import annotations

from typing import Any, Dict, List, Tuple

import torch


class Boxes:
    """(n, 4) boxes [x0, y0, x1, y1] (subset of detectron2.structures.Boxes)."""

    def __init__(self, tensor: torch.Tensor):
        self.tensor = tensor.reshape(-1, 4).to(torch.float32)

    def __len__(self) -> int:
        return int(self.tensor.shape[0])

    def __getitem__(self, item) -> "Boxes":
        t = self.tensor[item]
        return Boxes(t.reshape(1, 4) if t.dim() == 1 else t)

    def to(self, *args, **kwargs) -> "Boxes":
        return Boxes(self.tensor.to(*args, **kwargs))

    def get_centers(self) -> torch.Tensor:
        return (self.tensor[:, :2] + self.tensor[:, 2:]) / 2

    def scale(self, sx: float, sy: float) -> None:
        self.tensor[:, 0::2] *= sx
        self.tensor[:, 1::2] *= sy

    def clip(self, size: Tuple[int, int]) -> None:
        h, w = size
        self.tensor[:, 0::2].clamp_(min=0, max=w)
        self.tensor[:, 1::2].clamp_(min=0, max=h)

    def nonempty(self, threshold: float = 0.0) -> torch.Tensor:
        wh = self.tensor[:, 2:] - self.tensor[:, :2]
        return (wh[:, 0] > threshold) & (wh[:, 1] > threshold)

    @staticmethod
    def cat(boxes: List["Boxes"]) -> "Boxes":
        return Boxes(torch.cat([b.tensor for b in boxes], dim=0))


class Instances:
    """Per-image detections: a dict of equally long fields + image_size (height, width)."""

    def __init__(self, image_size: Tuple[int, int], **fields: Any):
        object.__setattr__(self, '_image_size', (int(image_size[0]), int(image_size[1])))
        object.__setattr__(self, '_fields', {})
        for k, v in fields.items():
            self.set(k, v)

    @property
    def image_size(self) -> Tuple[int, int]:
        return self._image_size

    def __setattr__(self, name: str, val: Any) -> None:
        if name.startswith('_'):
            object.__setattr__(self, name, val)
        else:
            self.set(name, val)

    def __getattr__(self, name: str) -> Any:
        fields = object.__getattribute__(self, '_fields')
        if name not in fields:
            raise AttributeError(f"Cannot find field '{name}' in the given Instances!")
        return fields[name]

    def set(self, name: str, value: Any) -> None:
        if self._fields:
            assert len(self) == len(value), f'Adding a field of length {len(value)} to Instances of length {len(self)}'
        self._fields[name] = value

    def has(self, name: str) -> bool:
        return name in self._fields

    def get(self, name: str) -> Any:
        return self._fields[name]

    def get_fields(self) -> Dict[str, Any]:
        return self._fields

    def remove(self, name: str) -> None:
        del self._fields[name]

    def __len__(self) -> int:
        for v in self._fields.values():
            return len(v)
        return 0

    def to(self, *args, **kwargs) -> "Instances":
        out = Instances(self._image_size)
        for k, v in self._fields.items():
            out.set(k, v.to(*args, **kwargs) if hasattr(v, 'to') else v)
        return out

    def __getitem__(self, item) -> "Instances":
        if isinstance(item, int):
            if item >= len(self) or item < -len(self):
                raise IndexError('Instances index out of range!')
            item = slice(item, None, len(self))
        out = Instances(self._image_size)
        for k, v in self._fields.items():
            out.set(k, v[item])
        return out

    @staticmethod
    def cat(instance_lists: List["Instances"]) -> "Instances":
        assert len(instance_lists) > 0
        if len(instance_lists) == 1:
            return instance_lists[0]
        out = Instances(instance_lists[0].image_size)
        for k in instance_lists[0]._fields:
            vals = [i.get(k) for i in instance_lists]
            if isinstance(vals[0], torch.Tensor):
                out.set(k, torch.cat(vals, dim=0))
            elif hasattr(type(vals[0]), 'cat'):
                out.set(k, type(vals[0]).cat(vals))
            else:
                out.set(k, sum(vals, []))
        return out

    def __repr__(self) -> str:
        return f'Instances(num_instances={len(self)}, image_size={self._image_size}, fields={list(self._fields)})'

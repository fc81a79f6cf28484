"""The reference's plugin protocol (news_rec_utils/pipeline.py:8-90), minus the joblib step cache
(storage is out of scope): components exchange a `context_dict`."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Iterable, Optional


def check_req_keys(required_keys: set, context_dict: dict) -> None:
    for key in required_keys:
        assert key in context_dict, f"Required Key {key} is not present in context_dict"


class PipelineComponent(ABC):
    required_keys: set = set()
    train_required_keys: set = set()

    @abstractmethod
    def transform(self, context_dict: dict) -> dict:
        ...

    def train(self, context_dict: dict, val_context_dict: Optional[dict] = None) -> None:
        return None


class Pipeline:
    def __init__(self, name: str, steps: Iterable[tuple[str, PipelineComponent]]):
        self.name = name
        self._steps = list(steps)

    def transform(self, context_dict: dict, val_context_dict: Optional[dict] = None):
        for _, comp in self._steps:
            context_dict = comp.transform(context_dict)
            if val_context_dict:
                val_context_dict = comp.transform(val_context_dict)
        return context_dict, val_context_dict

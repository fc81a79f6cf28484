"""Drop-in `FinalAttentionComponent` (reference: news_rec_utils/components.py:980-1027).

Only `transform` is in scope (scripts/eval.py's path); training stays with the reference."""
from __future__ import annotations

from pathlib import Path
from typing import Any, Optional

from .data_model_helper import get_final_second_attention_score
from .modeling_utils import get_final_attention_model
from .pipeline import PipelineComponent, check_req_keys


class FinalAttentionComponent(PipelineComponent):
    required_keys = {
        "news_embeddings",
        "impression_rev_ind_array",
        "impression_len_list",
        "history_rev_ind_array",
        "history_len_list",
        "history_bool",
    }
    train_required_keys = required_keys | {"labels"}

    def __init__(self, attention_model_path: Optional[Path] = None, attention_model=None, precision=None, **_unused):
        self.attention_model = attention_model if attention_model is not None else \
            get_final_attention_model(attention_model_path)
        self.precision = precision

    def transform(self, context_dict: dict[str, Any]) -> dict[str, Any]:
        check_req_keys(self.required_keys, context_dict)
        out = context_dict.copy()
        out.update(get_final_second_attention_score(
            out["history_rev_ind_array"][0], out["history_len_list"], out["impression_rev_ind_array"][0],
            out["impression_len_list"], out["news_embeddings"], out["history_bool"], self.attention_model,
            precision=self.precision))
        return out

    def train(self, context_dict, val_context_dict=None) -> None:
        raise NotImplementedError("training is out of scope for the B200 hot path (use the reference trainer)")

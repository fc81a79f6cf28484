"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Loads the UNMODIFIED reference (`/root/reference/src/news_rec_utils`) in THIS
container so that `oracle/make_golden.py` can mint golden vectors from the
reference's own code and `tests/` can pin `oracle/oracle.py` against it when
the tree is present.  `/root/reference` does not exist on the GPU box: nothing
that runs there may import this module (tests that use it skip when the path
is absent).

Shims -- none of them touch arithmetic:
  * transformers 5.x moved three names that the reference only uses as type
    hints (data_utils.py:18-22, modeling_utils.py:16);
  * `azure.storage.blob` (components.py:7, trainer.py:18) is absent -> stub;
  * the OOM-probing batch-size finder (batch_size_finder.py:103-149) cannot
    terminate on CPU -> constant; DataLoader workers -> 0 for determinism.
"""
from __future__ import annotations

import os
import sys
import types

_INSTALLED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")  # oracle/build_ref.py


def _reference_src() -> str:
    """The reference tree itself in the build container; on the GPU box the unmodified copy that
    oracle/build_ref.py pip-installed into oracle/_ref (git-ignored, travels with gpurun)."""
    env = os.environ.get("NRB_REFERENCE_SRC")
    if env:
        return env
    if os.path.isdir("/root/reference/src/news_rec_utils"):
        return "/root/reference/src"
    return _INSTALLED


REFERENCE_SRC = _reference_src()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "news_rec_utils"))


def reference_origin() -> str:
    return "oracle/_ref (pip-installed unmodified reference)" if REFERENCE_SRC == _INSTALLED else REFERENCE_SRC


_loaded = None


def load_reference(batch_size: int = 64, device: str | None = None):
    """Import the reference package; returns a namespace of its modules.

    `device`: "cpu" / "cuda" overrides the reference's DEVICE constant (config.py:19 picks cuda when it is
    available) in every module that imported it -- a run-time placement choice, no arithmetic."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_SRC}")
    sys.dont_write_bytecode = True  # the tree is read-only

    import transformers
    import transformers.tokenization_utils as tu
    import transformers.tokenization_utils_fast as tuf

    for mod in (tu, tuf):
        for name in ("PreTrainedTokenizer", "PreTrainedTokenizerFast", "BatchEncoding"):
            if not hasattr(mod, name):
                setattr(mod, name, getattr(transformers, name))

    if "azure.storage.blob" not in sys.modules:
        azure = types.ModuleType("azure")
        storage = types.ModuleType("azure.storage")
        blob = types.ModuleType("azure.storage.blob")

        class _Stub:  # pragma: no cover - never instantiated
            def __init__(self, *a, **k):
                raise RuntimeError("azure stub")

        blob.ContainerClient = _Stub
        blob.BlobClient = _Stub
        blob.BlobServiceClient = _Stub
        azure.storage = storage
        storage.blob = blob
        sys.modules.setdefault("azure", azure)
        sys.modules.setdefault("azure.storage", storage)
        sys.modules.setdefault("azure.storage.blob", blob)
    if "dotenv" not in sys.modules:
        try:
            import dotenv  # noqa: F401
        except Exception:
            dotenv = types.ModuleType("dotenv")
            dotenv.load_dotenv = lambda *a, **k: False
            sys.modules["dotenv"] = dotenv

    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)

    import news_rec_utils.config as config
    import news_rec_utils.latent_attention as latent_attention
    import news_rec_utils.evaluation as evaluation
    import news_rec_utils.pipeline as pipeline
    import news_rec_utils.data_utils as data_utils
    import news_rec_utils.modeling_utils as modeling_utils
    import news_rec_utils.data_model_helper as data_model_helper
    import news_rec_utils.components as components

    data_model_helper.get_attention_inference_batch_size = lambda model: 2 * batch_size
    data_model_helper.NUM_WORKERS = 0
    if device is not None:
        import torch

        dev = torch.device(device)
        for mod in (config, data_utils, modeling_utils, data_model_helper, components):
            if hasattr(mod, "DEVICE"):
                mod.DEVICE = dev

    ns = types.SimpleNamespace(
        config=config,
        latent_attention=latent_attention,
        evaluation=evaluation,
        pipeline=pipeline,
        data_utils=data_utils,
        modeling_utils=modeling_utils,
        data_model_helper=data_model_helper,
        components=components,
    )
    _loaded = ns
    return ns


def make_reference_latent_model(ref, dim: int, num_latents: int, seed: int = 1234):
    """Construct the reference LatentAttentionModel at (dim, num_latents).

    Dims come from module globals read at construction time
    (latent_attention.py:91-113); `latents` is replaced for L != 64.
    """
    import torch

    la = ref.latent_attention
    old = (la.REDUCED_DIM, la.EMBEDDING_DIM)
    la.REDUCED_DIM = dim
    la.EMBEDDING_DIM = dim
    try:
        torch.manual_seed(seed)
        model = la.LatentAttentionModel()
        if num_latents != model.latents.shape[0]:
            model.latents = torch.nn.Parameter(torch.randn(num_latents, dim))
    finally:
        la.REDUCED_DIM, la.EMBEDDING_DIM = old
    return model.eval()


def make_reference_final_attention(ref, dim: int, hidden: int, seed: int = 1234):
    import torch

    torch.manual_seed(seed)
    return ref.modeling_utils.FinalAttention(reduced_dim=dim, hidden_dim=hidden).eval()

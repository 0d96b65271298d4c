"""zsaac_b200 — B200-native related-caption retrieval for zero-shot audio captioning.

The package directory is `zero-shot-aac_b200/` (not an importable name); `import zsaac_b200`
at the repo root (zsaac_b200.py) registers it under the module name `zsaac_b200`.

Public surface (mirrors XinMing0411/zero-shot-AAC for the one hot path):
  data_handing.embeddings_related_generator          load_data / process_data / save_data_to_hdf5 / main
  data_handing.embeddings_related_generator_wavcaps  same, list of input files
  utils.sound_effect_choice                           top-k label retrieval
  retrieval.related_topk / retrieval.RelatedBank      batched scores + indices (new, low level)
  sharded.ShardedRelatedBank                          bank row-sharded over the GPUs of one box
"""
from . import _abi
from ._abi import ZsaacError, load_library
from .retrieval import RelatedBank, related_topk, clear_bank_cache

__all__ = ["RelatedBank", "related_topk", "clear_bank_cache", "ZsaacError", "load_library", "_abi"]
__version__ = "0.1.0"

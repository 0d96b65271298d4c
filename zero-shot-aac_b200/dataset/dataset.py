"""`collate` with the label retrieval batched in it — mirror of dataset/dataset.py:632-647.

In the reference every `Dataset.__getitem__` (dataset/dataset.py:365, :445, :530, :600) calls

    sound_effects_index = sound_effect_choice(prefix, self.sound_effect_embeddings, self.sound_effect_num).squeeze(0)
    selected_labels = [self.sound_effect_labels[i].lower() for i in list(sound_effects_index)]
    hard_prompt = parse_entities(self.tokenizer, selected_labels, self.mask_probability)

once per sample, inside the DataLoader workers (train_prompt_multilingual.py:60-61,
predict_mistralai_multilingual.py:90) — a [1, 1024] x [1024, 527] CPU matmul + softmax + topk per
item.  A forked worker cannot initialise CUDA, and this library has no CPU path, so the retrieval
moves to where the batch is assembled: `__getitem__` returns its tuple WITHOUT the two trailing
elements (hard_prompt, len(hard_prompt)), and `collate_with_sound_effects` — running in the main
process — does ONE batched `sound_effect_choice` for the B prefixes (one kernel launch for up to
64 samples), then the reference's own per-sample label lookup / `parse_entities` and its
`padding_captions`.  The returned tuple is exactly what the reference's `collate` returns.

`parse_entities` and `padding_captions` (utils.py:178-208: prompt text assembly, out of scope
here) are passed in by the caller, so this module imports nothing from the reference.
Use with `DataLoader(..., collate_fn=functools.partial(collate_with_sound_effects, ...))`; with
`num_workers > 0` the workers still run `__getitem__` (tokenisation, prefix lookup) in parallel.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import torch

from ..utils import sound_effect_choice


def attach_sound_effects(prefix: torch.Tensor, sound_effect_embeddings: torch.Tensor,
                         sound_effect_labels: Sequence[str], sound_effect_num: int, tokenizer,
                         mask_probability, parse_entities: Callable) -> List[torch.Tensor]:
    """hard_prompt of every sample of a batch: prefix [B, 1, d] or [B, d] -> list of B token
    tensors, each what the reference's __getitem__ builds for that sample (:365-368)."""
    index = sound_effect_choice(prefix, sound_effect_embeddings, sound_effect_num)   # [B, (1,) k] CPU
    index = index.reshape(-1, index.shape[-1])
    hard_prompts = []
    for sound_effects_index in index:
        selected_labels = [sound_effect_labels[i].lower() for i in list(sound_effects_index)]
        hard_prompts.append(parse_entities(tokenizer, selected_labels, mask_probability))
    return hard_prompts


def collate_with_sound_effects(batch, *, sound_effect_embeddings: torch.Tensor,
                               sound_effect_labels: Sequence[str], sound_effect_num: int, tokenizer,
                               parse_entities: Callable, padding_captions: Callable,
                               mask_probability=0):
    """Reference collate (dataset/dataset.py:632-647) for samples that carry no hard prompt yet.

    batch items: (tokens, mask, prefix) from the training datasets, or (audio_id, prefix) from the
    evaluation dataset.  Returns (tokens, mask, prefix, padding_hard_prompt, hard_prompts_masks)
    or (audio_id, prefix, padding_hard_prompt, hard_prompts_masks), like the reference."""
    training = len(batch[0]) == 3
    if training:
        tokens, mask, prefix = zip(*batch)
        tokens = torch.stack(tokens, dim=0)
        mask = torch.stack(mask)
    else:
        audio_id, prefix = zip(*batch)
    prefix = torch.stack(prefix)
    hard_prompt = attach_sound_effects(prefix, sound_effect_embeddings, sound_effect_labels,
                                       sound_effect_num, tokenizer, mask_probability, parse_entities)
    hard_prompt_length = [len(h) for h in hard_prompt]
    padding_hard_prompt, hard_prompts_masks = padding_captions(hard_prompt, hard_prompt_length)
    if training:
        return tokens, mask, prefix, padding_hard_prompt, hard_prompts_masks
    return audio_id, prefix, padding_hard_prompt, hard_prompts_masks


def read_related_records(data_path, caption_words=None) -> list:
    """The record reader of the reference's datasets, for the stream the generator scripts write.

    Reference: dataset/dataset.py:64-78 (ClapDataset: `pickle.load` until EOFError, a list object
    is spliced in, a dict is appended if its caption has 8-20 words -> caption_words=(8, 20)) and
    :401-417 (the hard-prompt datasets: every dict is kept -> caption_words=None).  `data_path`
    is a list of files like the reference's, or one path.  Same records as the reference's loop;
    the tensors inside are rebuilt by the direct parser of torch's per-tensor storage stream that
    load_data uses (related_pipeline._FastTensorUnpickler: ~3x faster than `pickle.load`, which
    runs a full torch.load per tensor; anything unusual goes through torch's own loader).  Host
    only — needs neither CUDA nor the native library."""
    import pickle
    from ..related_pipeline import _FastTensorUnpickler
    paths = [data_path] if isinstance(data_path, (str, bytes)) else list(data_path)
    all_data: list = []
    for dp in paths:
        with open(dp, "rb") as f:
            while True:
                try:
                    item = _FastTensorUnpickler(f).load()
                except EOFError:
                    break
                if type(item) is list:
                    all_data.extend(item)
                elif caption_words is None:
                    all_data.append(item)
                else:
                    n_words = len(item["caption"].split())
                    if caption_words[0] <= n_words <= caption_words[1]:
                        all_data.append(item)
    return all_data

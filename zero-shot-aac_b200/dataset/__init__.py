"""Mirror of the two pieces of the reference's dataset package that touch the hot path: the
per-sample label retrieval (`sound_effect_choice`) moved from `__getitem__` into `collate`, and
the reader loop for the record stream the generator scripts write."""
from .dataset import attach_sound_effects, collate_with_sound_effects, read_related_records  # noqa: F401

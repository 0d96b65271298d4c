"""Mirror of the one piece of the reference's dataset package that touches the hot path: the
per-sample label retrieval (`sound_effect_choice`) moved from `__getitem__` into `collate`."""
from .dataset import attach_sound_effects, collate_with_sound_effects  # noqa: F401

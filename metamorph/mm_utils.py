"""Alias of the reference module path `metamorph.mm_utils` (the helpers the hot path's callers use)."""
from metamorph_b200.mm_utils import tokenizer_image_token  # noqa: F401

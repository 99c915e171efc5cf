"""Constants of the reference (metamorph/constants.py:13-19 and the literals hard-coded in
metamorph_arch.py:317 / metamorph_llama.py:502), kept verbatim so the drop-in honours them."""
IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200
DEFAULT_IMAGE_TOKEN = "<image>"
DEFAULT_IMAGE_PATCH_TOKEN = "<im_patch>"
DEFAULT_IM_START_TOKEN = "<image_start>"
DEFAULT_IM_END_TOKEN = "<image_end>"
IMAGE_PLACEHOLDER = "<image-placeholder>"

IMAGE_START_TOKEN_ID = 128256   # metamorph_arch.py:317, metamorph_llama.py:502
IMAGE_END_TOKEN_ID = 128257
EOS_TOKEN_IDS = (128001, 128009)
VISION_FEATURE_DIM = 1152       # metamorph_llama.py:255, siglip_encoder.py:124

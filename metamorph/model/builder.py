from metamorph_b200.model.builder import load_pretrained_model  # noqa: F401

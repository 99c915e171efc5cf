from metamorph_b200.inference.load_metamorph import load_metamorph, load_metamorph_model  # noqa: F401

from metamorph_b200.model.metamorph_arch import *  # noqa: F401,F403
from metamorph_b200.model.metamorph_arch import LlavaMetaForCausalLM, MetaMorphMetaForCausalLM, MetaMorphMetaModel  # noqa: F401

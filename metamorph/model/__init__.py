from metamorph_b200.model import (LlavaMetaForCausalLM, MetaMorphConfig, MetaMorphLlamaForCausalLM,  # noqa: F401
                                  MetaMorphLlamaModel, MetaMorphMetaForCausalLM, MetaMorphMetaModel)
from metamorph_b200.model import metamorph_arch  # noqa: F401

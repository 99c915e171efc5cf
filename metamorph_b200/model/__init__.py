from .metamorph_arch import LlavaMetaForCausalLM, MetaMorphMetaForCausalLM, MetaMorphMetaModel
from .metamorph_llama import MetaMorphConfig, MetaMorphLlamaForCausalLM, MetaMorphLlamaModel

__all__ = ["MetaMorphConfig", "MetaMorphLlamaForCausalLM", "MetaMorphLlamaModel", "MetaMorphMetaModel",
           "MetaMorphMetaForCausalLM", "LlavaMetaForCausalLM"]

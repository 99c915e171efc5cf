"""Host-side mirror of `metamorph/model/metamorph_arch.py` for the hot path.

Same class / method names and argument meaning as the reference:
  MetaMorphMetaModel            (metamorph_arch.py:21)   builds tower + mm_projector (+ dead vision_proj)
  MetaMorphMetaForCausalLM      (metamorph_arch.py:131)  encode_images / encode_imagesembed /
                                                         prepare_inputs_labels_for_multimodal /
                                                         initialize_vision_tokenizer
The per-sample Python loop of the reference is replaced by `build_interleave_plan` (bit-exact index
logic on the host) + one CUDA gather kernel; see model/interleave_plan.py and csrc/interleave.cu.
"""
from __future__ import annotations

import re
from abc import ABC, abstractmethod

import torch
import torch.nn as nn

from .. import ops
from ..constants import DEFAULT_IM_END_TOKEN, DEFAULT_IM_START_TOKEN, DEFAULT_IMAGE_PATCH_TOKEN
from .interleave_plan import build_interleave_plan
from .layers import KernelLinear, MlpGelu
from .siglip_tower import build_vision_tower


class IdentityMap(nn.Module):
    def forward(self, x, *args, **kwargs):
        return x

    @property
    def config(self):
        return {"mm_projector_type": "identity"}


def build_vision_projector(config, delay_load=False, dtype=torch.bfloat16, device=None, **kwargs):
    """multimodal_projector/builder.py:39-64. In scope: mlp2x_gelu (every script), linear, identity."""
    projector_type = getattr(config, "mm_projector_type", "linear")
    if projector_type == "linear":
        return KernelLinear(config.mm_hidden_size, config.hidden_size, True, dtype, device)
    m = re.match(r"^mlp(\d+)x_gelu$", projector_type)
    if m:
        if int(m.group(1)) != 2:
            raise NotImplementedError("only mlp2x_gelu is used by the reference scripts (in scope)")
        return MlpGelu(config.mm_hidden_size, config.hidden_size, config.hidden_size, dtype, device)
    if projector_type == "identity":
        return IdentityMap()
    raise ValueError(f"Unknown projector type: {projector_type}")


class MetaMorphMetaModel:
    """Mixin for the LLaMA model class (metamorph_arch.py:21-96)."""

    def _init_vision(self, config, vision_delay_load=True, dtype=torch.bfloat16, device=None):
        if hasattr(config, "mm_vision_tower"):
            self.vision_tower = build_vision_tower(config, delay_load=vision_delay_load)
            if not vision_delay_load and device is not None:
                self.vision_tower.to(device)
            self.mm_projector = build_vision_projector(config, dtype=dtype, device=device)
            self.vision_proj = KernelLinear(4096, config.hidden_size, True, dtype, device)  # dead param (:31)

    def get_vision_tower(self):
        vision_tower = getattr(self, "vision_tower", None)
        if type(vision_tower) is list:
            vision_tower = vision_tower[0]
        return vision_tower

    def initialize_vision_modules(self, model_args, fsdp=None):
        """metamorph_arch.py:45-96: build tower/projector from model_args, copy knobs onto config,
        optionally load a stage-1 `mm_projector.bin`."""
        vision_tower = model_args.vision_tower
        self.config.mm_vision_tower = vision_tower
        for k in ("mm_vision_select_layer", "mm_vision_select_feature", "image_token_reduction",
                  "num_image_tokens", "freeze_vision", "normalize_vision", "apply_softmax", "vision_coef"):
            if hasattr(model_args, k):
                setattr(self.config, k, getattr(model_args, k))
        dev = self.embed_tokens.weight.device
        dt = self.embed_tokens.weight.dtype
        if self.get_vision_tower() is None:
            tower = build_vision_tower(model_args, delay_load=True)
            self.vision_tower = tower
        else:
            tower = self.get_vision_tower()
        if not tower.is_loaded:
            tower.load_model(device=dev, dtype=dt)
        self.config.use_mm_proj = True
        self.config.mm_projector_type = getattr(model_args, "mm_projector_type", "linear")
        self.config.mm_hidden_size = tower.hidden_size
        self.config.mm_patch_merge_type = getattr(model_args, "mm_patch_merge_type", "flat")
        if getattr(self, "mm_projector", None) is None:
            self.mm_projector = build_vision_projector(self.config, dtype=dt, device=dev)
            self.vision_proj = KernelLinear(4096, self.config.hidden_size, True, dt, dev)
        else:
            for p in self.mm_projector.parameters():
                p.requires_grad = True
        pretrain = getattr(model_args, "pretrain_mm_mlp_adapter", None)
        if pretrain is not None:
            w = torch.load(pretrain, map_location="cpu")
            sub = {k.split("mm_projector.")[1]: v for k, v in w.items() if "mm_projector" in k}
            self.mm_projector.load_state_dict(sub)


class MetaMorphMetaForCausalLM(ABC):
    @abstractmethod
    def get_model(self):
        pass

    def get_vision_tower(self):
        return self.get_model().get_vision_tower()

    def encode_images(self, images, return_prob=False):
        """metamorph_arch.py:140-164 -> (projected features [N,n,H], detached tower features [N,n,C])."""
        if return_prob:
            raise NotImplementedError("return_prob (mlpsoftmax projector) is unused by the reference scripts")
        image_features = self.get_model().get_vision_tower()(images)
        ar = self.get_model().mm_projector(image_features)
        return ar, image_features.detach().clone()

    def encode_imagesembed(self, image_features, return_prob=False):
        ar = self.get_model().mm_projector(image_features)
        return ar, image_features.detach().clone()

    def plan_inputs(self, input_ids, attention_mask, labels, num_images):
        cfg = self.config
        return build_interleave_plan(
            input_ids, attention_mask, labels, num_images,
            image_len=self.get_vision_tower().image_token_len,
            tokenizer_model_max_length=getattr(cfg, "tokenizer_model_max_length", None),
            padding_side=getattr(cfg, "tokenizer_padding_side", "right"))

    def prepare_inputs_labels_for_multimodal(self, input_ids, position_ids, attention_mask,
                                             past_key_values, labels, images, image_sizes=None,
                                             image_embeds=None, use_vision=True):
        """metamorph_arch.py:177-425. Returns the same 8-tuple:
        (None, position_ids, attention_mask, past_key_values, inputs_embeds, labels,
         image_positions, target_features)."""
        vision_tower = self.get_vision_tower()
        if image_embeds is None:
            if vision_tower is None or images is None or input_ids.shape[1] == 1 or not use_vision:
                if not use_vision:
                    input_ids = input_ids[input_ids != -200]
                return input_ids, position_ids, attention_mask, past_key_values, None, labels, None, None
            if type(images) is list or images.ndim == 5:
                raise NotImplementedError("list / 5-D (anyres) image inputs are unused by the reference scripts")
            image_features, target_features = self.encode_images(images)
        else:
            image_features, target_features = self.encode_imagesembed(image_embeds)
        dev = image_features.device
        n_img, n_tok, H = image_features.shape
        plan = build_interleave_plan(
            input_ids, attention_mask, labels, n_img, n_tok,
            getattr(self.config, "tokenizer_model_max_length", None),
            getattr(self.config, "tokenizer_padding_side", "right"))
        embeds = ops.interleave_gather(self.get_model().embed_tokens.weight.data,
                                       image_features.reshape(n_img * n_tok, H).contiguous(),
                                       plan.row_map.reshape(-1).to(dev))
        embeds = embeds.view(plan.batch, plan.seq_len, H)
        out_dev = input_ids.device
        new_labels = None if labels is None else plan.labels.to(labels.device)
        new_mask = None if attention_mask is None else plan.attention_mask.to(device=attention_mask.device,
                                                                             dtype=attention_mask.dtype)
        new_pos = None if position_ids is None else plan.position_ids.to(out_dev)
        if len(plan.target_image_idx) == n_img:
            tgt = target_features
        else:
            tgt = target_features[torch.tensor(plan.target_image_idx, dtype=torch.long, device=dev)]
        self._last_plan = plan
        return None, new_pos, new_mask, past_key_values, embeds, new_labels, plan.image_positions.to(out_dev), tgt

    def initialize_vision_tokenizer(self, model_args, tokenizer):
        """metamorph_arch.py:427-469 (token surgery for <image_start>/<image_end>)."""
        if model_args.mm_use_im_patch_token:
            tokenizer.add_tokens([DEFAULT_IMAGE_PATCH_TOKEN], special_tokens=True)
            self.resize_token_embeddings(len(tokenizer))
        if model_args.mm_use_im_start_end:
            num_new_tokens = tokenizer.add_tokens([DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN], special_tokens=True)
            self.resize_token_embeddings(len(tokenizer))
            if num_new_tokens > 0:
                ie = self.get_input_embeddings().weight.data
                oe = self.get_output_embeddings().weight.data
                ie[-num_new_tokens:] = ie[:-num_new_tokens].float().mean(dim=0, keepdim=True).to(ie.dtype)
                oe[-num_new_tokens:] = oe[:-num_new_tokens].float().mean(dim=0, keepdim=True).to(oe.dtype)
            if model_args.tune_mm_mlp_adapter:
                for p in self.get_input_embeddings().parameters():
                    p.requires_grad = True
                for p in self.get_output_embeddings().parameters():
                    p.requires_grad = False
            if model_args.pretrain_mm_mlp_adapter:
                w = torch.load(model_args.pretrain_mm_mlp_adapter, map_location="cpu")
                etw = w["model.embed_tokens.weight"]
                assert num_new_tokens == 2
                ie = self.get_input_embeddings().weight.data
                if ie.shape == etw.shape:
                    ie[-num_new_tokens:] = etw[-num_new_tokens:]
                elif etw.shape[0] == num_new_tokens:
                    ie[-num_new_tokens:] = etw
                else:
                    raise ValueError(f"Unexpected embed_tokens_weight shape. Pretrained: {etw.shape}. "
                                     f"Current: {ie.shape}. Numer of new tokens: {num_new_tokens}.")
        elif model_args.mm_use_im_patch_token:
            if model_args.tune_mm_mlp_adapter:
                for p in self.get_input_embeddings().parameters():
                    p.requires_grad = False
                for p in self.get_output_embeddings().parameters():
                    p.requires_grad = False


# north-star spelling (BASELINE.json) of the same mixin
LlavaMetaForCausalLM = MetaMorphMetaForCausalLM

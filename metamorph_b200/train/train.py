"""`train()` entry of the hot path (mirrors metamorph/train/train.py:1375 — same argument dataclasses and
field names :44-113 — but drives `TrainEngine` instead of HF Trainer + DeepSpeed).

Scope: the reference's dataset / tokenisation / prompt-template code is host-side and
out of scope; this entry consumes batches that already follow the collator contract
(`input_ids` with -200 placeholders, `labels`, `attention_mask`, `images`; train.py:1258-1284) from
  * `--data_path synthetic[:B,T]`  seeded synthetic interleaved batches (bench / smoke), or
  * a user supplied `data_module_factory(tokenizer, data_args) -> iterable of batches`.
Launch exactly like the reference (one process per GPU): `torchrun --nproc-per-node N -m metamorph_b200.train.train ...`.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Callable, Iterable, Optional

import torch
import torch.distributed as dist


@dataclass
class ModelArguments:
    model_name_or_path: Optional[str] = field(default="facebook/opt-125m")
    version: Optional[str] = field(default="v0")
    freeze_backbone: bool = field(default=False)
    tune_mm_mlp_adapter: bool = field(default=False)
    vision_tower: Optional[str] = field(default=None)
    mm_vision_select_layer: Optional[int] = field(default=-1)
    pretrain_mm_mlp_adapter: Optional[str] = field(default=None)
    mm_projector_type: Optional[str] = field(default="linear")
    mm_use_im_start_end: bool = field(default=False)
    mm_use_im_patch_token: bool = field(default=True)
    mm_patch_merge_type: Optional[str] = field(default="flat")
    mm_vision_select_feature: Optional[str] = field(default="patch")
    num_image_tokens: int = field(default=256)
    image_token_reduction: str = field(default="interpolation")
    use_vision_ar: bool = field(default=True)
    vision_coef: float = field(default=1.0)
    vision_head_type: str = field(default="mlp")
    normalize_vision: bool = field(default=True)
    apply_softmax: bool = field(default=False)
    freeze_vision: bool = field(default=True)


@dataclass
class DataArguments:
    data_path: str = field(default="synthetic")
    lazy_preprocess: bool = False
    is_multimodal: bool = True
    image_folder: Optional[str] = field(default=None)
    image_aspect_ratio: str = "square"


@dataclass
class TrainingArguments:
    output_dir: str = field(default="./checkpoints")
    optim: str = field(default="adamw_torch")
    model_max_length: int = field(default=4096)
    per_device_train_batch_size: int = field(default=4)
    gradient_accumulation_steps: int = field(default=1)
    learning_rate: float = field(default=6.93e-5)
    weight_decay: float = field(default=0.0)
    warmup_ratio: float = field(default=0.03)
    lr_scheduler_type: str = field(default="cosine")
    max_steps: int = field(default=100)
    max_grad_norm: Optional[float] = field(default=None)
    bf16: bool = field(default=True)
    logging_steps: int = field(default=1)
    seed: int = field(default=42)
    gradient_checkpointing: bool = field(default=True)  # accepted for CLI compatibility; selective recompute is built in
    save_gu_layers: int = field(default=32)
    save_steps: int = field(default=0)          # > 0: write output_dir/checkpoint-<step> every save_steps steps
    pack_sequences: bool = field(default=False)  # lay a batch's samples end to end instead of padding
    pack_len: Optional[int] = field(default=None)  # row length of the packed layout (default: the batch maximum)


def rank0_print(*args):
    if not dist.is_initialized() or dist.get_rank() == 0:
        print(*args, flush=True)


def _synthetic_batches(spec: str, training_args, model_args, rank: int):
    from .. import synthetic
    B, T = training_args.per_device_train_batch_size, training_args.model_max_length
    if ":" in spec:
        B, T = (int(x) for x in spec.split(":")[1].split(","))
    step = 0
    while True:
        yield synthetic.train_batch(B, T, image_tokens=model_args.num_image_tokens, seed=1234 + 7919 * step + 1000 * rank)
        step += 1


def train(attn_implementation=None, data_module_factory: Optional[Callable[..., Iterable[dict]]] = None,
          argv=None):
    from transformers import HfArgumentParser
    from ..engine.trainer import TrainEngine
    from ..model import MetaMorphConfig, MetaMorphLlamaForCausalLM
    from .. import synthetic

    parser = HfArgumentParser((ModelArguments, DataArguments, TrainingArguments))
    model_args, data_args, training_args = parser.parse_args_into_dataclasses(args=argv)
    torch.manual_seed(training_args.seed)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    rank = dist.get_rank() if world > 1 else 0

    if os.path.isdir(model_args.model_name_or_path):
        model = MetaMorphLlamaForCausalLM.from_pretrained(
            model_args.model_name_or_path, torch_dtype=torch.bfloat16, device=dev,
            use_vision_ar=model_args.use_vision_ar, vision_coef=model_args.vision_coef,
            vision_head=model_args.vision_head_type, normalize_vision=model_args.normalize_vision,
            apply_softmax=model_args.apply_softmax)
    else:  # no checkpoint on disk (offline box): random init at LLaMA-3-8B dims
        cfg = synthetic.make_config(num_image_tokens=model_args.num_image_tokens,
                                    max_len=training_args.model_max_length)
        model = MetaMorphLlamaForCausalLM(cfg, use_vision_ar=model_args.use_vision_ar,
                                          vision_coef=model_args.vision_coef, vision_head=model_args.vision_head_type,
                                          normalize_vision=model_args.normalize_vision, device=dev)
    model.config.use_cache = False
    if model_args.vision_tower is not None and model_args.vision_tower != "None":
        # train.py:1498-1502: always (re-)initialise the vision modules from model_args — this is also what loads a
        # stage-1 `--pretrain_mm_mlp_adapter` (metamorph_arch.py:91-96), whether or not the config already names a tower
        model.get_model().initialize_vision_modules(model_args=model_args)
    tower = model.get_vision_tower()
    if tower is not None and not tower.is_loaded:
        tower.load_model(device=dev)          # checkpoint tensors > pretrained SigLIP > (synthetic configs only) random
    from ..constants import IMAGE_END_TOKEN_ID
    if model_args.mm_use_im_start_end and model.config.vocab_size <= IMAGE_END_TOKEN_ID:
        # the reference grows the vocabulary by <image_start>/<image_end> through initialize_vision_tokenizer
        # (train.py:1545, needs the tokenizer); without a tokenizer the checkpoint must already contain both rows
        raise ValueError(f"vocab_size {model.config.vocab_size} has no rows for <image_start>/<image_end> "
                         f"({IMAGE_END_TOKEN_ID - 1}, {IMAGE_END_TOKEN_ID}): run initialize_vision_tokenizer(model_args, tokenizer) "
                         "on the model first")
    for k in ("mm_use_im_start_end", "mm_use_im_patch_token", "num_image_tokens", "image_token_reduction",
              "normalize_vision", "freeze_vision", "vision_coef", "vision_head_type"):
        setattr(model.config, k, getattr(model_args, k))
    model.config.tokenizer_model_max_length = training_args.model_max_length
    model.config.tokenizer_padding_side = "right"
    if tower is not None:
        for p in tower.parameters():
            p.requires_grad = False
    if model_args.tune_mm_mlp_adapter:  # stage 1: only the projector trains (train.py:1516-1519)
        for p in model.parameters():
            p.requires_grad = False
        for p in model.get_model().mm_projector.parameters():
            p.requires_grad = True

    engine = TrainEngine(model, lr=training_args.learning_rate, weight_decay=training_args.weight_decay,
                         max_grad_norm=training_args.max_grad_norm, total_steps=training_args.max_steps,
                         warmup_ratio=training_args.warmup_ratio, n_save_gu_layers=training_args.save_gu_layers,
                         pack_sequences=training_args.pack_sequences, pack_len=training_args.pack_len)
    if data_module_factory is not None:
        batches = iter(data_module_factory(None, data_args))
    elif data_args.data_path.startswith("synthetic"):
        batches = _synthetic_batches(data_args.data_path, training_args, model_args, rank)
    else:
        raise NotImplementedError(
            "the reference's JSONL/image dataset pipeline (train.py LazySupervisedDataset) is host-side code "
            "outside the hot path; pass data_module_factory=... or --data_path synthetic")
    model.train()
    from .. import checkpoint
    start = 0
    resume = checkpoint.latest_checkpoint(training_args.output_dir) if training_args.output_dir else None
    if resume is not None:                                    # train.py:1592-1595: checkpoint-* present -> resume
        start = checkpoint.load_training_checkpoint(engine, resume)
        rank0_print(f"resumed from {resume} at step {start}")
        for _ in range(start * accum):                        # keep the data stream aligned with the step counter
            next(batches)
    elif training_args.output_dir and os.path.isdir(training_args.output_dir) and any(
            d.startswith("checkpoint-") for d in os.listdir(training_args.output_dir)):
        rank0_print(f"{training_args.output_dir} holds checkpoint-* folders without trainer_state.json (stage-1 projector "
                    "artefacts are not resumable): starting from step 0")
    for step in range(start, training_args.max_steps):
        # one optimizer step = `gradient_accumulation_steps` micro-batches of per_device_train_batch_size samples
        out = engine.step([next(batches) for _ in range(accum)] if accum > 1 else next(batches))
        if (step + 1) % training_args.logging_steps == 0:
            vals = torch.cat([out["loss"].reshape(1), out["loss_language"].reshape(1), out["loss_image_ar"].reshape(1)])
            if world > 1:
                dist.all_reduce(vals)
                vals /= world
            l, ll_, li = vals.tolist()
            model.loss_language, model.loss_image_ar = ll_, li
            rank0_print(f"step {step + 1}: loss {l:.4f} loss_language {ll_:.4f} loss_image_ar {li:.4f} lr {engine.current_lr:.3e}")
        if training_args.save_steps > 0 and (step + 1) % training_args.save_steps == 0 and training_args.output_dir:
            torch.cuda.synchronize()
            if model_args.tune_mm_mlp_adapter:                # metamorph_trainer.py:273-291
                if rank == 0:
                    checkpoint.save_mm_projector_checkpoint(model, training_args.output_dir, step + 1,
                                                            use_im_start_end=model_args.mm_use_im_start_end)
            else:                                             # every rank: the optimizer state is sharded over ranks
                checkpoint.save_training_checkpoint(engine, training_args.output_dir)
            if world > 1:
                dist.barrier()
    if rank == 0 and training_args.output_dir:
        torch.cuda.synchronize()
        if model_args.tune_mm_mlp_adapter:                    # safe_save_model_for_hf_trainer, train.py:189-207
            checkpoint.save_mm_projector(model, training_args.output_dir,
                                         use_im_start_end=model_args.mm_use_im_start_end)
        else:
            model.save_pretrained(training_args.output_dir)
    return model


if __name__ == "__main__":
    train()

"""Checkpoint I/O in the reference's formats, straight from the fused device layout.

What the reference writes / reads (all through HF Trainer / torch.save):
  * full model: `trainer._save(output_dir, state_dict=cpu_state_dict)` (train.py:213-222) -> HF `save_pretrained`:
    `model.safetensors`, or `model-0000i-of-0000n.safetensors` + `model.safetensors.index.json` when the weights
    exceed `max_shard_size` (5 GB in transformers 4.45), keys = the HF names of SURVEY §8b's state-dict contract;
  * stage-1 artefact: only the `mm_projector` (+ `embed_tokens` with `use_im_start_end`) parameters as
    `mm_projector.bin`, or `<parent>/mm_projector/checkpoint-N.bin` when the output folder is a `checkpoint-N`
    (train.py:189-207, metamorph_trainer.py:273-291); loaded back with `load_state_dict(strict=False)`
    (builder.py:82-84) or into the bare projector (metamorph_arch.py:92-96);
  * resume: `checkpoint-*` folders under `output_dir` trigger `trainer.train(resume_from_checkpoint=True)`
    (train.py:1592-1595).
Here the weights live fused on the device (packed QKV, interleaved gate/up): `iter_reference_state` un-fuses ONE
tensor at a time and hands it over on the CPU, so saving an 8 B model never needs a second device copy nor the whole
state dict in host memory at once (shards are flushed as they fill). The training checkpoint adds the optimizer's
fp32 master / m / v (this framework's own layout, like DeepSpeed's optimizer shards) and the step counter; loading
restores the engine state bit for bit and the resumed run tracks the uninterrupted one (the backward's fp32 atomics
make two runs agree to rounding, not bitwise) — tests/test_trainer_gpu.py.
"""
from __future__ import annotations

import json
import os
import re
from typing import Dict, Iterator, List, Optional, Tuple

import torch

WEIGHTS_NAME = "model.safetensors"
WEIGHTS_INDEX_NAME = "model.safetensors.index.json"
SHARD_PATTERN = "model-{:05d}-of-{:05d}.safetensors"
OPT_PATTERN = "optimizer-{:05d}-of-{:05d}.safetensors"
OPT_INDEX_NAME = "optimizer.safetensors.index.json"
TRAINER_STATE_NAME = "trainer_state.json"
PREFIX_CHECKPOINT_DIR = "checkpoint"      # transformers.trainer_utils.PREFIX_CHECKPOINT_DIR


def parse_size(size) -> int:
    """'5GB' / '200MB' / int -> bytes (decimal units, as huggingface_hub's parse_size_to_int)."""
    if isinstance(size, int):
        return size
    m = re.fullmatch(r"\s*(\d+(?:\.\d+)?)\s*([KMGT]?I?B)\s*", size.upper())
    if not m:
        raise ValueError(f"cannot parse size {size!r}")
    num, unit = float(m.group(1)), m.group(2)
    mult = {"B": 1, "KB": 10 ** 3, "MB": 10 ** 6, "GB": 10 ** 9, "TB": 10 ** 12,
            "KIB": 2 ** 10, "MIB": 2 ** 20, "GIB": 2 ** 30, "TIB": 2 ** 40}[unit]
    return int(num * mult)


# --------------------------------------------------------------------------------------------- weights
def iter_reference_state(model) -> Iterator[Tuple[str, torch.Tensor]]:
    """(HF name, CPU tensor) pairs in `model.state_dict()` order, one tensor materialised at a time."""
    from .engine.packing import deinterleave_gate_up
    from .model.metamorph_llama import dims_from_config
    d = dims_from_config(model.config)
    qw, kw = d.n_heads * d.head_dim, d.n_kv_heads * d.head_dim
    for k, v in torch.nn.Module.state_dict(model).items():
        v = v.detach()
        if k.endswith("self_attn.qkv_proj.weight"):
            base = k[:-len("qkv_proj.weight")]
            yield base + "q_proj.weight", v[:qw].cpu().contiguous()
            yield base + "k_proj.weight", v[qw:qw + kw].cpu().contiguous()
            yield base + "v_proj.weight", v[qw + kw:].cpu().contiguous()
        elif k.endswith("mlp.gate_up_proj.weight"):
            base = k[:-len("gate_up_proj.weight")]
            g, u = deinterleave_gate_up(v)
            yield base + "gate_proj.weight", g.cpu().contiguous()
            yield base + "up_proj.weight", u.cpu().contiguous()
        else:
            yield k, v.cpu().contiguous()
    tower = model.get_vision_tower()
    if tower is not None and tower.is_loaded:
        for k, v in tower.vision_tower._extra_state_tensors.items():
            yield "model.vision_tower.vision_tower." + k, v.detach().cpu().contiguous()


def _plan_shards(sizes: List[Tuple[str, int]], max_bytes: int) -> List[List[str]]:
    """Greedy in iteration order, like huggingface_hub.split_torch_state_dict_into_shards: a tensor that would make the
    current shard exceed max_bytes opens a new one; a single tensor larger than the limit gets its own shard."""
    shards, cur, cur_bytes = [], [], 0
    for name, nb in sizes:
        if cur and cur_bytes + nb > max_bytes:
            shards.append(cur)
            cur, cur_bytes = [], 0
        cur.append(name)
        cur_bytes += nb
    if cur:
        shards.append(cur)
    return shards


def _write_sharded(pairs: Iterator[Tuple[str, torch.Tensor]], sizes: List[Tuple[str, int]], out_dir: str,
                   max_shard_size, single_name: str, pattern: str, index_name: str) -> Dict[str, str]:
    from safetensors.torch import save_file
    os.makedirs(out_dir, exist_ok=True)
    plan = _plan_shards(sizes, parse_size(max_shard_size))
    n = len(plan)
    files = [single_name] if n == 1 else [pattern.format(i + 1, n) for i in range(n)]
    where = {name: files[i] for i, names in enumerate(plan) for name in names}
    weight_map: Dict[str, str] = {}
    cur_file, cur = None, {}

    def flush():
        if cur_file is not None and cur:
            save_file(cur, os.path.join(out_dir, cur_file), metadata={"format": "pt"})

    for name, t in pairs:
        f = where[name]
        if f != cur_file:
            flush()
            cur_file, cur = f, {}
        cur[name] = t
        weight_map[name] = f
    flush()
    if n > 1:
        index = {"metadata": {"total_size": sum(nb for _, nb in sizes)}, "weight_map": weight_map}
        with open(os.path.join(out_dir, index_name), "w") as fh:
            json.dump(index, fh, indent=2, sort_keys=True)
            fh.write("\n")
    return weight_map


def reference_state_sizes(model) -> List[Tuple[str, int]]:
    """(HF name, bytes) in save order without materialising anything."""
    from .model.metamorph_llama import dims_from_config
    d = dims_from_config(model.config)
    qw, kw = d.n_heads * d.head_dim, d.n_kv_heads * d.head_dim
    out = []
    for k, v in torch.nn.Module.state_dict(model).items():
        es = v.element_size()
        if k.endswith("self_attn.qkv_proj.weight"):
            base, cols = k[:-len("qkv_proj.weight")], v.shape[1]
            out += [(base + "q_proj.weight", qw * cols * es), (base + "k_proj.weight", kw * cols * es),
                    (base + "v_proj.weight", kw * cols * es)]
        elif k.endswith("mlp.gate_up_proj.weight"):
            base = k[:-len("gate_up_proj.weight")]
            half = v.numel() // 2 * es
            out += [(base + "gate_proj.weight", half), (base + "up_proj.weight", half)]
        else:
            out.append((k, v.numel() * es))
    tower = model.get_vision_tower()
    if tower is not None and tower.is_loaded:
        for k, v in tower.vision_tower._extra_state_tensors.items():
            out.append(("model.vision_tower.vision_tower." + k, v.numel() * v.element_size()))
    return out


def save_model(model, out_dir: str, max_shard_size="5GB") -> Dict[str, str]:
    """HF-format weights + config.json, as `trainer._save` / `save_pretrained(safe_serialization=True)` leave them.
    Returns the weight map (tensor name -> file)."""
    os.makedirs(out_dir, exist_ok=True)
    model.config.save_pretrained(out_dir)
    return _write_sharded(iter_reference_state(model), reference_state_sizes(model), out_dir, max_shard_size,
                          WEIGHTS_NAME, SHARD_PATTERN, WEIGHTS_INDEX_NAME)


def load_model_state(path: str) -> Dict[str, torch.Tensor]:
    """All HF-named tensors of a checkpoint directory: sharded / single safetensors, or pytorch_model*.bin."""
    from safetensors.torch import load_file
    idx = os.path.join(path, WEIGHTS_INDEX_NAME)
    sd: Dict[str, torch.Tensor] = {}
    if os.path.exists(idx):
        with open(idx) as fh:
            files = sorted(set(json.load(fh)["weight_map"].values()))
        for f in files:
            sd.update(load_file(os.path.join(path, f)))
        return sd
    for f in sorted(os.listdir(path)):
        fp = os.path.join(path, f)
        if f.endswith(".safetensors") and not f.startswith("optimizer"):
            sd.update(load_file(fp))
        elif f.startswith("pytorch_model") and f.endswith(".bin"):
            sd.update(torch.load(fp, map_location="cpu"))
    return sd


# --------------------------------------------------------------------------------------------- stage-1 artefact
def mm_projector_state(model, use_im_start_end: bool = False, extra_keys=()) -> Dict[str, torch.Tensor]:
    """get_mm_adapter_state_maybe_zero_3(named_parameters, keys_to_match) of train.py:163-166,192-196."""
    keys = ["mm_projector", *extra_keys]
    if use_im_start_end:
        keys += ["embed_tokens", "embed_in"]
    return {k: p.detach().cpu() for k, p in model.named_parameters() if any(m in k for m in keys)}


def save_mm_projector(model, output_dir: str, use_im_start_end: bool = False) -> str:
    """safe_save_model_for_hf_trainer's tune_mm_mlp_adapter branch (train.py:189-207): config.json into output_dir,
    weights into `<parent>/mm_projector/checkpoint-N.bin` for a checkpoint folder, else `output_dir/mm_projector.bin`."""
    weights = mm_projector_state(model, use_im_start_end)
    os.makedirs(output_dir, exist_ok=True)
    model.config.save_pretrained(output_dir)
    current = output_dir.rstrip("/").split("/")[-1]
    if current.startswith(PREFIX_CHECKPOINT_DIR + "-"):
        folder = os.path.join(os.path.dirname(output_dir.rstrip("/")), "mm_projector")
        os.makedirs(folder, exist_ok=True)
        dst = os.path.join(folder, f"{current}.bin")
    else:
        dst = os.path.join(output_dir, "mm_projector.bin")
    torch.save(weights, dst)
    return dst


def save_mm_projector_checkpoint(model, run_dir: str, global_step: int, use_im_start_end: bool = False) -> str:
    """MetaMorphTrainer._save_checkpoint's tune_mm_mlp_adapter branch (metamorph_trainer.py:273-291): the periodic
    stage-1 checkpoint is `run_dir/checkpoint-<step>/{config.json, mm_projector.bin}`."""
    out = os.path.join(run_dir, f"{PREFIX_CHECKPOINT_DIR}-{global_step}")
    os.makedirs(out, exist_ok=True)
    weights = mm_projector_state(model, use_im_start_end, extra_keys=("vision_resampler",))
    model.config.save_pretrained(out)
    dst = os.path.join(out, "mm_projector.bin")
    torch.save(weights, dst)
    return dst


def load_mm_projector(model, path: str):
    """builder.py:82-84: `model.load_state_dict(torch.load('mm_projector.bin'), strict=False)` (dtype follows the model)."""
    if os.path.isdir(path):
        path = os.path.join(path, "mm_projector.bin")
    weights = torch.load(path, map_location="cpu")
    dtype = next(model.parameters()).dtype
    res = model.load_state_dict({k: v.to(dtype) for k, v in weights.items()}, strict=False)
    # the artefact holds only the projector (+ embeddings with mm_use_im_start_end): what matters is that every tensor in
    # it found its parameter and that the projector itself is complete
    bad = [k for k in res.unexpected_keys]
    lacking = [k for k in res.missing_keys if k.startswith("model.mm_projector.")]
    if bad or lacking:
        raise RuntimeError(f"load_mm_projector({path}): unexpected tensors {bad[:4]}, projector tensors missing {lacking[:4]}")
    return res


# --------------------------------------------------------------------------------------------- training checkpoints
def _opt_items(engine) -> Iterator[Tuple[str, torch.Tensor, torch.Tensor, torch.Tensor]]:
    """(key, master, exp_avg, exp_avg_sq). With the whole state on the rank (1 GPU, or shard_optimizer=False) the keys are
    this framework's fused parameter names; with the state sharded over ranks (engine/trainer.py) a bucket contributes
    this rank's slice under `<bucket>::slice<r>of<w>` and the small replicated tensors are written by rank 0 only."""
    if engine.shard_world == 1:
        for name, st in engine.opt.items():
            yield name, st.p32, st.m, st.v
        return
    for b in list(engine.layer_buckets) + list(engine.big_buckets.values()):
        yield f"{b.name}::slice{engine.shard_rank}of{engine.shard_world}", b.p32, b.m, b.v
    if engine.rank == 0:
        for name, st in engine.small_state.items():
            yield name, st.p32, st.m, st.v


def _opt_pairs(engine) -> Iterator[Tuple[str, torch.Tensor]]:
    for name, p32, m, v in _opt_items(engine):
        yield name + "::master", p32.detach().cpu().contiguous()
        yield name + "::exp_avg", m.detach().cpu().contiguous()
        yield name + "::exp_avg_sq", v.detach().cpu().contiguous()


def _rank_opt_file(rank: int, world: int) -> str:
    return f"optimizer-rank{rank:05d}-of-{world:05d}.safetensors"


def save_training_checkpoint(engine, output_dir: str, max_shard_size="5GB") -> str:
    """`output_dir/checkpoint-<step>/`: HF-format weights (as the reference's Trainer checkpoints) + optimizer state
    (fp32 master weights, exp_avg, exp_avg_sq under this framework's fused parameter names) + trainer_state.json.
    With a sharded optimizer EVERY rank must call this: rank 0 writes the weights and trainer_state.json, each rank its
    own `optimizer-rank*-of-*.safetensors` (as DeepSpeed ZeRO writes one optimizer file per rank); resuming then needs
    the same world size."""
    ckpt = os.path.join(output_dir, f"{PREFIX_CHECKPOINT_DIR}-{engine.step_count}")
    if engine.rank == 0:
        save_model(engine.model, ckpt, max_shard_size)
    os.makedirs(ckpt, exist_ok=True)
    if engine.shard_world == 1:
        if engine.rank == 0:
            sizes = [(f"{n}::{part}", p32.numel() * 4) for n, p32, _, _ in _opt_items(engine)
                     for part in ("master", "exp_avg", "exp_avg_sq")]
            _write_sharded(_opt_pairs(engine), sizes, ckpt, max_shard_size, "optimizer.safetensors", OPT_PATTERN,
                           OPT_INDEX_NAME)
    else:
        from safetensors.torch import save_file
        save_file(dict(_opt_pairs(engine)), os.path.join(ckpt, _rank_opt_file(engine.shard_rank, engine.shard_world)))
    if engine.rank == 0:
        state = {"global_step": engine.step_count, "max_steps": engine.total_steps, "learning_rate": engine.lr,
                 "warmup_ratio": engine.warmup_ratio, "constant_lr": engine.constant_lr, "betas": list(engine.betas),
                 "eps": engine.eps, "weight_decay": engine.wd, "max_grad_norm": engine.max_grad_norm,
                 "world_size": engine.world, "optimizer_shards": engine.shard_world,
                 "gradient_accumulation_steps": engine.accum}
        with open(os.path.join(ckpt, TRAINER_STATE_NAME), "w") as fh:
            json.dump(state, fh, indent=2, sort_keys=True)
            fh.write("\n")
    return ckpt


def latest_checkpoint(output_dir: str) -> Optional[str]:
    """The newest `checkpoint-N` under output_dir (train.py:1592 globs for them; HF resumes from the highest N)."""
    if not os.path.isdir(output_dir):
        return None
    best, best_n = None, -1
    for f in os.listdir(output_dir):
        m = re.fullmatch(PREFIX_CHECKPOINT_DIR + r"-(\d+)", f)
        if m and os.path.isdir(os.path.join(output_dir, f)) and int(m.group(1)) > best_n and \
                os.path.exists(os.path.join(output_dir, f, TRAINER_STATE_NAME)):
            best, best_n = os.path.join(output_dir, f), int(m.group(1))
    return best


def load_training_checkpoint(engine, ckpt_dir: str) -> int:
    """Restores weights (bf16 compute copies), fp32 master weights, both moments and the step counter in place.
    Returns the restored global step."""
    from safetensors.torch import load_file
    with open(os.path.join(ckpt_dir, TRAINER_STATE_NAME)) as fh:
        state = json.load(fh)
    res = engine.model.load_state_dict(load_model_state(ckpt_dir), strict=False)
    engine.model.check_loaded_keys(res, f"load_training_checkpoint({ckpt_dir})")
    trainable_missing = [k for k in res.missing_keys if k in set(engine.trainable_names)]
    if trainable_missing:
        raise RuntimeError(f"load_training_checkpoint({ckpt_dir}): trainable tensors missing: {trainable_missing[:4]}")
    dests = {}
    for name, p32, m, v in _opt_items(engine):
        dests[name] = {"master": p32, "exp_avg": m, "exp_avg_sq": v}
    if engine.shard_world > 1:
        if int(state.get("optimizer_shards", 1)) != engine.shard_world:
            raise RuntimeError(f"{ckpt_dir} holds {state.get('optimizer_shards', 1)} optimizer shard(s); resuming a sharded "
                               f"optimizer needs the same world size (now {engine.shard_world})")
        for name, st in engine.small_state.items():            # replicated tensors live in rank 0's file
            dests[name] = {"master": st.p32, "exp_avg": st.m, "exp_avg_sq": st.v}
        files = sorted({_rank_opt_file(engine.shard_rank, engine.shard_world), _rank_opt_file(0, engine.shard_world)})
    else:
        idx = os.path.join(ckpt_dir, OPT_INDEX_NAME)
        if os.path.exists(idx):
            with open(idx) as fh:
                files = sorted(set(json.load(fh)["weight_map"].values()))
        else:
            files = ["optimizer.safetensors"]
    seen = set()
    for f in files:
        for key, t in load_file(os.path.join(ckpt_dir, f)).items():
            name, part = key.rsplit("::", 1)
            d = dests.get(name)
            if d is None:
                if engine.shard_world > 1 and "::slice" in name:
                    continue                                    # another rank's slice in rank 0's file
                raise KeyError(f"optimizer state for unknown parameter {name!r}")
            dst = d[part]
            dst.copy_(t.to(dst.device).view_as(dst))
            seen.add((name, part))
    missing = [n for n in dests if (n, "master") not in seen]
    if missing:
        raise KeyError(f"optimizer state missing for {missing[:3]}... ({len(missing)} parameters)")
    engine.refresh_compute_copies()      # the bf16 compute copies are the rounded masters, as after every step
    engine.step_count = int(state["global_step"])
    tower = engine.model.get_vision_tower()
    if tower is not None and tower.is_loaded:
        tower.vision_tower.invalidate_packed()
    return engine.step_count

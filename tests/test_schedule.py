"""CPU: the engine's learning-rate schedule against HF's own scheduler driven the way HF Trainer drives it
(`--lr_scheduler_type cosine --warmup_ratio 0.03`, scripts/*.sh): optimizer.step() then lr_scheduler.step()."""
import math

import torch


def test_lr_used_at_every_optimizer_step_matches_hf_trainer():
    from transformers import get_cosine_schedule_with_warmup
    from metamorph_b200.engine.trainer import cosine_lr
    base, total, ratio = 6.93e-5, 200, 0.03
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=base)
    sched = get_cosine_schedule_with_warmup(opt, num_warmup_steps=math.ceil(total * ratio), num_training_steps=total)
    for k in range(1, total + 1):                       # k = 1-based optimizer step == TrainEngine.step_count in step()
        hf_lr = opt.param_groups[0]["lr"]               # the lr HF's optimizer.step() uses now
        ours = cosine_lr(max(k - 1, 0), total, base, ratio)
        assert abs(hf_lr - ours) <= 1e-12 * base + 1e-18, (k, hf_lr, ours)
        opt.step()
        sched.step()


def test_engine_current_lr_indexing():
    from metamorph_b200.engine.trainer import TrainEngine, cosine_lr

    class _E:                                            # the property only needs these attributes
        constant_lr, lr, total_steps, warmup_ratio = False, 1e-3, 10, 0.2
    e = _E()
    for k, want in [(0, 0.0), (1, 0.0), (2, cosine_lr(1, 10, 1e-3, 0.2)), (3, 1e-3)]:
        e.step_count = k
        assert TrainEngine.current_lr.fget(e) == want

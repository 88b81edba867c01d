"""Learning-rate schedule of the reference's training recipe (README.md:147-151: `--lr_scheduler_type cosine
--warmup_ratio 0.03`, HF Trainer -> transformers.get_cosine_schedule_with_warmup, un-vendored third-party code whose
published formula is restated here): linear warm-up over ceil(warmup_ratio * total) steps, then half a cosine to 0.
Host arithmetic only; FineTuner.set_lr applies the factor to both the base and the mm_projector learning rate."""
import math


def warmup_steps(num_training_steps, warmup_ratio=0.03):
    """TrainingArguments.get_warmup_steps: ceil(total * ratio)."""
    return math.ceil(num_training_steps * warmup_ratio)


def cosine_with_warmup(step, num_training_steps, num_warmup_steps, num_cycles=0.5):
    """Multiplier of the base lr for optimizer step `step` (0-based: the value used BY step `step`, i.e. after `step`
    scheduler.step() calls)."""
    if step < num_warmup_steps:
        return float(step) / float(max(1, num_warmup_steps))
    progress = float(step - num_warmup_steps) / float(max(1, num_training_steps - num_warmup_steps))
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(num_cycles) * 2.0 * progress)))

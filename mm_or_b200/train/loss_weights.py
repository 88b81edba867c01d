"""Class weights of the fine-tune loss.

Reference: train/train.py:1310-1322 turns the token-frequency file (`--token_weight_path`, e.g.
data/llava_samples/train_token_freqs_7b_50perm.json: 96 sentence pieces of the scene-graph vocabulary with their
counts) into weights 1 / (ln f + 1), every other token gets min(weight) / 100; LLaVATrainer.compute_loss
(train/llava_trainer.py:143-151) scatters them into a vector over the tokenizer's vocabulary that becomes
nn.CrossEntropyLoss(weight=...). The vector is what FineTuner(vocab_weight=...) / b200_weighted_ce consume.
"""
import json
import math

import torch


def vocab_weight_from_frequencies(token_frequencies, vocab):
    """token_frequencies: {piece: count} (or a path to the JSON); vocab: {piece: id} as tokenizer.get_vocab() returns.
    -> fp32 tensor (len(vocab),). A piece missing from the vocabulary raises KeyError like the reference."""
    if isinstance(token_frequencies, str):
        with open(token_frequencies) as f:
            token_frequencies = json.load(f)
    weights = {k: 1.0 / (math.log(v) + 1.0) for k, v in token_frequencies.items()}     # train.py:1319
    extra = min(weights.values()) / 100                                                # train.py:1320-1321
    out = torch.ones(len(vocab)) * extra                                               # llava_trainer.py:146
    for k, v in weights.items():
        out[vocab[k]] = v                                                              # llava_trainer.py:148-149
    return out

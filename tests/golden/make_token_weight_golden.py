"""Records the class-weight vector the reference builds from its own token-frequency file
(data/llava_samples/train_token_freqs_7b_50perm.json, the fixture SURVEY.md 4 lists) with the arithmetic of
train/train.py:1316-1321 (numpy log, as there) and train/llava_trainer.py:146-149, over a synthetic vocabulary that
places the 96 pieces at known ids (the Llama tokenizer is not available offline). -> tests/golden/token_weights.pt

Run in the build container only:   python tests/golden/make_token_weight_golden.py
"""
import json
import os

import numpy as np
import torch

SRC = "/root/reference/scene_graph_generation/data/llava_samples/train_token_freqs_7b_50perm.json"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "token_weights.pt")


def main():
    with open(SRC) as f:
        token_frequencies = json.load(f)
    token_weights = {k: 1 / (np.log(v) + 1) for k, v in token_frequencies.items()}
    min_weight = min(token_weights.values())
    extra_token_weight = min_weight / 100
    pieces = sorted(token_frequencies)
    vocab = {p: 7 + 3 * i for i, p in enumerate(pieces)}                  # known ids inside a 32000-entry vocabulary
    for i in range(32000):
        vocab.setdefault(f"<filler{i}>", i) if i not in vocab.values() else None
    ids_used = set(vocab.values())
    assert len(vocab) == 32000 and ids_used == set(range(32000))
    vocab_weight = torch.ones(len(vocab)) * extra_token_weight
    for k, v in token_weights.items():
        vocab_weight[vocab[k]] = v
    torch.save({"frequencies": token_frequencies, "piece_ids": {p: vocab[p] for p in pieces},
                "vocab_weight_nonextra": {int(vocab[p]): float(vocab_weight[vocab[p]]) for p in pieces},
                "extra": float(vocab_weight[0]), "sum": float(vocab_weight.double().sum())}, OUT)
    print(len(pieces), "pieces; extra weight", extra_token_weight, "sum", float(vocab_weight.double().sum()))


if __name__ == "__main__":
    main()

"""Shared helpers for the parity tests."""
import torch

from oracle import mm2sg_oracle as O


def oracle_cfg(cfg):
    return O.cfg_from_llava(cfg)


def rel_err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def max_err(a, b):
    return (a.float().cpu() - b.float().cpu()).abs().max().item()


# bf16 tolerance used throughout the GPU parity tests: relative Frobenius error of a bf16 pipeline against the fp32
# oracle evaluated on the same bf16-rounded weights and inputs. One bf16 rounding is 2^-9 ~ 0.2 % per element; after
# the ~10 rounding points of a transformer block the accumulated relative error stays below 1 %, and below 2 % after
# the full small model (2 ViT + 2 BERT + 2 Llama blocks).
TOL_OP = 1e-2
TOL_STAGE = 2e-2
TOL_E2E = 3e-2

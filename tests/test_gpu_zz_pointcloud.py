"""GPU: the point-cloud branch (PointTransformerV3, csrc/ptv3.cu) through the C ABI against the oracle and the fixtures
recorded from the reference (same checks as the CPU emulator run, tests/ptv3_checks.py), and end to end through
`model.generate(..., pc=...)`. Runs last in the suite (file name) -- the newest kernels."""
import pytest
import torch

import golden_cases as gc
import ptv3_checks as C
from helpers import TOL_E2E, TOL_STAGE, oracle_cfg, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from mm_or_b200.model.point_transformer import PcOps
    torch.cuda.set_device(0)
    return PcOps.cuda("cuda:0")


@pytest.mark.parametrize("check", C.ALL, ids=lambda f: f.__name__[6:])
def test_pointcloud_operator(ops, check):
    from mm_or_b200 import _lib as L
    n0 = L.launch_count()
    check(ops)
    torch.cuda.synchronize()
    if getattr(check, "no_launch", False):
        assert L.launch_count() == n0, "a rejected call launched a kernel"
    else:
        assert L.launch_count() > n0, "no kernel of libb200mmor.so was launched"


@pytest.mark.parametrize("order", range(4))
def test_codes_match_reference_fixture(ops, order):
    C.check_codes_match_reference_fixture(ops, order)


def test_host_pointers_are_rejected():
    from mm_or_b200 import _lib as L
    from mm_or_b200.model.point_transformer import PcOps
    ops = PcOps.cuda("cuda:0")
    with pytest.raises(L.B200Error):
        ops.grid_coords(torch.zeros(4, 6), 0.01)        # a CPU tensor: no fallback


def test_generate_with_point_clouds_matches_oracle():
    """pc token in the multimodal pack: tokens (pooled, pc, audio, seg x 3) and prefill logits vs the oracle."""
    from mm_or_b200.model.llava_llama import LlavaLlamaForCausalLM
    from oracle import mm2sg_oracle as O
    from oracle import ptv3_oracle as P
    torch.set_grad_enabled(False)
    cfg = gc.small_config()
    sd = gc.small_weights(cfg)
    sd.update(P.synth_weights())
    sd = gc.bf16_round(sd)
    model = LlavaLlamaForCausalLM(cfg).load_state_dict(sd, device="cuda:0")
    model.config.tokenizer_padding_side = "left"
    case = gc.make_case(cfg, "extras_left")
    pcs = P.dedupe_clouds([P.synth_cloud(1300, seed=21), None, P.synth_cloud(300, seed=22, box=(16, 16, 3))])
    torch.manual_seed(77)
    out, lg = model.generate(case["input_ids"], images=case["images"], audio=case["audio"], segmasks=case["segmasks"],
                             pc=pcs, max_new_tokens=3, stop_on_eos=False, return_logits=True)
    ocfg = oracle_cfg(cfg)
    torch.manual_seed(77)
    ref = O.multimodal_prefill(sd, ocfg, case["input_ids"], case["attention_mask"], case["images"], audio=case["audio"],
                               segmasks=case["segmasks"], pc=pcs, padding_side="left")
    assert ref["visual"].shape[1] == 576 + 5
    toks, ref_lg = O.greedy_decode(sd, ocfg, ref["logits"][:, -1], ref["kv"], ref["mask"], 3, stop_on_eos=False)
    assert rel_err(lg, ref_lg) < TOL_E2E
    assert out.shape[1] == case["input_ids"].shape[1] + 3
    # the same batch was recorded from the reference's own forward (tests/golden/make_pc_golden.py)
    import os
    fx = torch.load(os.path.join(gc.GOLDEN_DIR, "pc_left.pt"))
    assert fx["shuffle_seed"] == 77 and [None if c is None else len(c) for c in pcs] == fx["n_points"]
    assert rel_err(lg[:, :3], fx["greedy_logits"][:, :3].float()) < TOL_E2E

"""CPU: the host-side pack planner (mm_or_b200/model/pack.py) against the oracle's restatement of
prepare_inputs_labels_for_multimodal (llava_arch.py:235-338), including the edge cases the reference handles:
ragged lengths, interior pads, left / right padding, truncation, text-only rows."""
import numpy as np
import pytest
import torch

from mm_or_b200.constants import IGNORE_INDEX, IMAGE_TOKEN_INDEX
from mm_or_b200.model.pack import PAD_ROW, VISUAL_BASE, plan_pack
from oracle import mm2sg_oracle as O


def random_batch(rng, B, Lt, with_labels, interior_pad=False, text_only_row=False):
    ids = torch.zeros(B, Lt, dtype=torch.long)
    for b in range(B):
        n = int(rng.integers(3, Lt + 1))
        row = torch.from_numpy(rng.integers(3, 500, n))
        if not (text_only_row and b == 0):
            row[int(rng.integers(0, n))] = IMAGE_TOKEN_INDEX
        if interior_pad and n > 4:
            row[2] = 0
        ids[b, Lt - n:] = row
    labels = None
    if with_labels:
        labels = ids.clone()
        labels[ids <= 0] = IGNORE_INDEX
    return ids, ids.ne(0), labels


@pytest.mark.parametrize("side", ["left", "right"])
@pytest.mark.parametrize("with_labels", [False, True])
@pytest.mark.parametrize("max_len", [None, 40])
def test_plan_matches_oracle(side, with_labels, max_len):
    rng = np.random.default_rng(0)
    for trial in range(8):
        ids, mask, labels = random_batch(rng, B=4, Lt=20, with_labels=with_labels, interior_pad=trial % 2 == 1,
                                         text_only_row=trial == 3)
        t_vis = int(rng.integers(1, 33))
        src, lab, m, pos = O.pack_plan(ids, mask, labels, t_vis, side, max_len)
        p = plan_pack(ids.numpy(), mask.numpy(), None if labels is None else labels.numpy(), t_vis, side, max_len)
        assert np.array_equal(p.src, src.numpy())
        assert np.array_equal(p.labels, lab.numpy())
        assert np.array_equal(p.mask, m.numpy())
        assert np.array_equal(p.pos, pos.numpy())
        # derived tables
        assert np.array_equal(p.lengths, m.sum(1).numpy())
        for b in range(ids.shape[0]):
            if side == "left" and p.lengths[b]:
                assert p.kv_start[b] == p.L - p.lengths[b]
            for j in range(t_vis):
                r = p.row_map[b * t_vis + j]
                where = np.where(p.src[b] == VISUAL_BASE - j)[0]
                if len(where):
                    assert r == b * p.L + where[0]
                else:
                    assert r == -1


def test_no_mask_no_labels():
    ids = torch.tensor([[5, IMAGE_TOKEN_INDEX, 6, 7]])
    p = plan_pack(ids.numpy(), None, None, 3)
    assert p.src.tolist() == [[5, -2, -3, -4, 6, 7]]
    assert p.labels.tolist() == [[IGNORE_INDEX] * 6]
    assert p.row_map.tolist() == [1, 2, 3]


def test_empty_row_and_truncation():
    ids = torch.tensor([[0, 0, 0, 0], [9, IMAGE_TOKEN_INDEX, 8, 7]])
    p = plan_pack(ids.numpy(), ids.ne(0).numpy(), None, 5, "left", max_len=4)
    assert p.L == 4 and p.lengths.tolist() == [0, 4]
    assert (p.src[0] == PAD_ROW).all()
    assert p.src[1].tolist() == [9, -2, -3, -4]              # visual tokens 3, 4 and the trailing text are truncated
    assert p.row_map.tolist()[5:] == [5, 6, 7, -1, -1]


def test_several_image_placeholders_match_reference_fixture():
    """24 batches whose rows hold 0-3 <image> placeholders through the REFERENCE's own function
    (tests/golden/make_pack_golden.py -> pack_multi_image_cases.pt; llava_arch.py:239,245,253-276): the k-th placeholder
    of the batch takes the k-th entry of `images` (one running index; a text-only row skips one), and asking for more
    blocks than there are raises IndexError like the reference. Planner and oracle reproduce source rows, the block of
    every visual position, labels, mask and position ids bit for bit."""
    import os
    recs = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pack_multi_image_cases.pt"))
    assert len(recs) == 24 and sum("raises" in r for r in recs) == 8
    for r in recs:
        labels = r["labels"]
        args = (r["ids"].numpy(), r["mask"].numpy(), None if labels is None else labels.numpy(), r["t_vis"], r["side"],
                r["max_len"])
        if "raises" in r:
            with pytest.raises(IndexError, match="out of bounds"):
                plan_pack(*args, n_blocks=r["n_blocks"])
            with pytest.raises(IndexError, match="out of bounds"):
                O.pack_plan(r["ids"], r["mask"], labels, r["t_vis"], r["side"], r["max_len"], n_blocks=r["n_blocks"])
            assert "out of bounds" in r["message"]
            continue
        p = plan_pack(*args, n_blocks=r["n_blocks"])
        src, lab, m, pos, blk = O.pack_plan(r["ids"], r["mask"], labels, r["t_vis"], r["side"], r["max_len"],
                                            n_blocks=r["n_blocks"], return_blocks=True)
        ref_mask, ref_blk, t_vis = r["out_mask"].numpy(), r["blocks"].numpy(), r["t_vis"]
        vis = p.vis_ids.reshape(p.src.shape)
        planner_blk = np.where(vis >= 0, vis // t_vis, -1)
        assert np.array_equal(planner_blk, ref_blk) and np.array_equal(blk.numpy(), ref_blk)
        assert np.array_equal(np.where(vis >= 0, vis % t_vis, -1), np.where(ref_blk >= 0, VISUAL_BASE - p.src, -1))
        for got_src, got_lab, got_mask, got_pos in ((p.src, p.labels, p.mask, p.pos),
                                                    (src.numpy(), lab.numpy(), m.numpy(), pos.numpy())):
            assert np.array_equal(got_mask, ref_mask)
            assert np.array_equal(got_src, r["src"].numpy())
            if r["out_labels"] is not None:
                assert np.array_equal(got_lab, r["out_labels"].numpy())
            assert np.array_equal(got_pos * ref_mask, r["out_pos"].numpy() * ref_mask)
        # row_map is the inverse of vis_ids
        flat = p.vis_ids
        for dst in np.where(flat >= 0)[0]:
            assert p.row_map[flat[dst]] == dst
        assert int((p.row_map >= 0).sum()) == int((flat >= 0).sum())


def test_two_images_one_row():
    ids = torch.tensor([[IMAGE_TOKEN_INDEX, 4, IMAGE_TOKEN_INDEX]])
    p = plan_pack(ids.numpy(), None, None, 2, n_blocks=2)
    assert p.src.tolist() == [[-2, -3, 4, -2, -3]]
    assert p.vis_ids.tolist() == [0, 1, -2, 2, 3] and p.row_map.tolist() == [0, 1, 3, 4]
    with pytest.raises(IndexError):                          # one block for two placeholders, as the reference
        plan_pack(ids.numpy(), None, None, 2)


def test_plan_matches_reference_fixture():
    """32 batches (left / right padding, labels or not, truncation, interior pads, a text-only row) through the
    REFERENCE's own prepare_inputs_labels_for_multimodal (tests/golden/make_pack_golden.py -> pack_cases.pt): the
    planner and the oracle reproduce source rows, labels, mask and position ids bit for bit."""
    import os
    recs = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pack_cases.pt"))
    assert len(recs) == 32
    for r in recs:
        labels = r["labels"]
        p = plan_pack(r["ids"].numpy(), r["mask"].numpy(), None if labels is None else labels.numpy(), r["t_vis"],
                      r["side"], r["max_len"])
        src, lab, m, pos = O.pack_plan(r["ids"], r["mask"], labels, r["t_vis"], r["side"], r["max_len"])
        ref_mask = r["out_mask"].numpy()
        assert p.src.shape == tuple(r["src"].shape), (r["side"], r["max_len"])
        for got_src, got_lab, got_mask, got_pos in ((p.src, p.labels, p.mask, p.pos),
                                                    (src.numpy(), lab.numpy(), m.numpy(), pos.numpy())):
            assert np.array_equal(got_mask, ref_mask)
            assert np.array_equal(got_src, r["src"].numpy())
            if r["out_labels"] is not None:                  # the reference returns None when no labels were passed
                assert np.array_equal(got_lab, r["out_labels"].numpy())
            # pad rows: the reference leaves 0 there (llava_arch.py:313,331-338)
            assert np.array_equal(got_pos * ref_mask, r["out_pos"].numpy() * ref_mask)


def test_plan_with_vis_descriptors_matches_reference_fixture():
    """24 batches with VIS_DESCRIPTOR placeholders and vis_descriptor_embs through the REFERENCE's own
    prepare_inputs_labels_for_multimodal (tests/golden/make_pack_golden.py -> pack_desc_cases.pt; llava_arch.py:243,
    253-294): placeholders before and after <image> (the visual tokens always take the FIRST split point), fewer
    descriptors than placeholders (one zero row each), surplus descriptors (ignored), multi-row descriptors, a text-only
    row (placeholders stay ordinary tokens), truncation, both padding sides, the bare-list form of a batch of one.
    Planner and oracle reproduce source rows, labels, mask and position ids bit for bit."""
    import os
    from mm_or_b200.model.pack import DESC_BASE
    recs = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pack_desc_cases.pt"))
    assert len(recs) == 24 and any(r["bare_list"] for r in recs)
    n_desc_rows = n_dummy = 0
    for r in recs:
        labels = r["labels"]
        p = plan_pack(r["ids"].numpy(), r["mask"].numpy(), None if labels is None else labels.numpy(), r["t_vis"],
                      r["side"], r["max_len"], desc_rows=r["desc_rows"])
        src, lab, m, pos = O.pack_plan(r["ids"], r["mask"], labels, r["t_vis"], r["side"], r["max_len"],
                                       desc_rows=r["desc_rows"])
        ref_mask, ref_src = r["out_mask"].numpy(), r["src"].numpy()
        assert p.src.shape == ref_src.shape
        for got_src, got_lab, got_mask, got_pos in ((p.src, p.labels, p.mask, p.pos),
                                                    (src.numpy(), lab.numpy(), m.numpy(), pos.numpy())):
            assert np.array_equal(got_mask, ref_mask)
            assert np.array_equal(got_src, ref_src)
            if r["out_labels"] is not None:
                assert np.array_equal(got_lab, r["out_labels"].numpy())
            assert np.array_equal(got_pos * ref_mask, r["out_pos"].numpy() * ref_mask)
        # desc_ids: the batch-wide table row behind every descriptor position, "untouched" (-2) elsewhere
        first = np.concatenate([[0], np.cumsum([sum(d) for d in r["desc_rows"]])])
        flat = ref_src.reshape(-1).astype(np.int64)
        sample = np.repeat(np.arange(ref_src.shape[0]), ref_src.shape[1])
        want = np.where(flat <= DESC_BASE, first[sample] + (DESC_BASE - flat), -2)
        assert np.array_equal(p.desc_ids, want)
        n_desc_rows += int((flat <= DESC_BASE).sum())
        n_dummy += int(((ref_src == -1) & ref_mask).sum())
    assert n_desc_rows > 20 and n_dummy > 3                    # the fixture exercises real and dummy descriptors

"""Host-side index arithmetic of the multimodal token pack.

Mirror of the control flow in LlavaMetaForCausalLM.prepare_inputs_labels_for_multimodal
(reference: LLaVA/llava/model/llava_arch.py:235-338): strip padding by the attention mask, split every row at its
<image> placeholders (-200), splice T_vis visual tokens in per placeholder (the k-th placeholder of the batch takes the
k-th block of image features -- the reference's running `cur_image_idx`, :239,263-266), IGNORE_INDEX labels over them, truncate to
tokenizer_model_max_length, pad left or right to the batch maximum, rebuild mask / position_ids. The reference does
this with ~40 tiny torch ops and .tolist() syncs per sample; here it is pure numpy on the host and produces the
index tables the device kernels consume (b200_projector_pack / b200_embed_rows), so no embedding row is ever
copied twice.
"""
import numpy as np

from ..constants import IGNORE_INDEX, IMAGE_TOKEN_INDEX, VIS_DESCRIPTOR_TOKEN_INDEX

PAD_ROW = -1        # zero embedding row
VISUAL_BASE = -2    # src == VISUAL_BASE - j  <=>  j-th visual token of the sample
DESC_BASE = -(1 << 20)   # src == DESC_BASE - r  <=>  r-th row of the sample's concatenated vis_descriptor_embs


class PackPlan:
    __slots__ = ("src", "labels", "mask", "pos", "lengths", "kv_start", "row_map", "L", "t_vis", "desc_ids", "vis_ids",
                 "n_blocks")


def descriptor_row_counts(vis_descriptor_embs, batch):
    """Row counts of the reference's vis_descriptor_embs argument (llava_arch.py:278-290): a list with one list of
    tensors per sample -- or, for a batch of one, the bare list of tensors (:279-280); a 1-D tensor is one row (:288).
    Returns (per-sample lists of tensors, per-sample lists of row counts)."""
    embs = vis_descriptor_embs
    if len(embs) > 0 and type(embs[0]) is not list:
        embs = [embs]
    if len(embs) < batch:
        raise IndexError(f"vis_descriptor_embs has {len(embs)} entries for a batch of {batch}")   # as :283 would
    return embs, [[1 if e.ndim == 1 else int(e.shape[0]) for e in per] for per in embs]


def plan_pack(input_ids, attention_mask, labels, t_vis, padding_side="right", max_len=None, desc_rows=None,
              n_blocks=None):
    """input_ids (B, Lt) int array with IMAGE_TOKEN_INDEX placeholders; attention_mask / labels optional.
    desc_rows: None, or per sample the row counts of its vis_descriptor_embs (descriptor_row_counts): every
    VIS_DESCRIPTOR_TOKEN_INDEX position of a row with an image is then replaced by the rows of the next descriptor
    (one zero row when the sample has fewer descriptors than placeholders, llava_arch.py:284-286) followed by the text
    up to the next placeholder (:278-294). With desc_rows=None that text is dropped, as in the reference.
    n_blocks: number of image-feature blocks (= len(images); default: one per row). The reference walks them with ONE
    running index over the whole batch (`cur_image_idx`, llava_arch.py:239): every <image> placeholder takes the next
    block, a text-only row skips one (:244-250), and running past the last block is its IndexError. With one
    placeholder per row (every MM2SG prompt) block b lands in row b; a prompt with several <image> placeholders takes
    several consecutive blocks (`images` then holds more view groups than there are rows).
    Returns a PackPlan:
      desc_ids (B * L,) int32 or None: row of the batch's concatenated descriptor table to copy to that packed
               position, -2 elsewhere (the "leave untouched" id of b200_embed_rows)
      src      (B, L) int32   >= 0 token id | PAD_ROW | VISUAL_BASE - j | DESC_BASE - r
      labels   (B, L) int64   IGNORE_INDEX on visual and pad rows
      mask     (B, L) bool
      pos      (B, L) int64   arange over real rows, 0 on pads
      lengths  (B,)   int32   real rows per sample
      kv_start (B,)   int32   first real row (left padding) else 0
      row_map  (n_blocks * t_vis,) int32  flat row (b * L + l) of visual token (block k, j), or -1 if truncated / absent
      vis_ids  (B * L,) int32  k * t_vis + j of the visual token at that packed position, -2 elsewhere (b200_embed_rows)
    """
    ids = np.asarray(input_ids)
    B = ids.shape[0]
    am = np.ones_like(ids, dtype=bool) if attention_mask is None else np.asarray(attention_mask).astype(bool)
    lab = np.full_like(ids, IGNORE_INDEX) if labels is None else np.asarray(labels)
    n_blocks = B if n_blocks is None else int(n_blocks)
    rows, rlabels, rblocks = [], [], []
    k = 0                                           # cur_image_idx (llava_arch.py:239)

    def take_block(b):
        nonlocal k
        if k >= n_blocks:                           # image_features[cur_image_idx] (:245,264) raises the same
            raise IndexError(f"index {k} is out of bounds for dimension 0 with size {n_blocks} (row {b} asks for "
                             "more image-feature blocks than `images` holds)")
        k += 1
        return k - 1

    for b in range(B):
        r_ids = ids[b][am[b]]                      # compaction: interior pads are dropped too (llava_arch.py:235)
        r_lab = lab[b][am[b]]
        n_img = int((r_ids == IMAGE_TOKEN_INDEX).sum())
        if n_img == 0:                              # text-only row (llava_arch.py:244-251): consumes a block index
            take_block(b)
            rows.append(r_ids.astype(np.int64))
            rlabels.append(r_lab)
            rblocks.append(np.full(len(r_ids), -1, dtype=np.int64))
            continue
        cuts = np.where((r_ids == IMAGE_TOKEN_INDEX) | (r_ids == VIS_DESCRIPTOR_TOKEN_INDEX))[0]
        bounds = [-1] + cuts.tolist() + [len(r_ids)]
        vis = VISUAL_BASE - np.arange(t_vis, dtype=np.int64)
        vlab = np.full(t_vis, IGNORE_INDEX, dtype=r_lab.dtype)
        parts, lparts, bparts = [], [], []
        # only the chunks up to the image count are emitted when vis_descriptor_embs is None (llava_arch.py:268-294).
        # NB the reference splits at <image> AND descriptor placeholders alike and then takes the first n_img + 1
        # chunks, inserting a feature block after each of the first n_img -- whatever kind of placeholder ended it.
        for i in range(n_img + 1):
            parts.append(r_ids[bounds[i] + 1:bounds[i + 1]].astype(np.int64))
            lparts.append(r_lab[bounds[i] + 1:bounds[i + 1]])
            bparts.append(np.full(len(parts[-1]), -1, dtype=np.int64))
            if i < n_img:
                parts.append(vis)
                lparts.append(vlab)
                bparts.append(np.full(t_vis, take_block(b), dtype=np.int64))
        if desc_rows is not None:
            n_desc = int((r_ids == VIS_DESCRIPTOR_TOKEN_INDEX).sum())
            first = np.concatenate([[0], np.cumsum(desc_rows[b])]).astype(np.int64)
            for j in range(n_desc):
                if j < len(desc_rows[b]):
                    drow = DESC_BASE - np.arange(first[j], first[j + 1], dtype=np.int64)
                else:
                    drow = np.full(1, PAD_ROW, dtype=np.int64)          # "Using dummy tensor": zeros(4096)
                parts.append(drow)
                lparts.append(np.full(len(drow), IGNORE_INDEX, dtype=r_lab.dtype))
                bparts.append(np.full(len(drow), -1, dtype=np.int64))
                parts.append(r_ids[bounds[n_img + j + 1] + 1:bounds[n_img + j + 2]].astype(np.int64))
                lparts.append(r_lab[bounds[n_img + j + 1] + 1:bounds[n_img + j + 2]])
                bparts.append(np.full(len(parts[-1]), -1, dtype=np.int64))
        rows.append(np.concatenate(parts))
        rlabels.append(np.concatenate(lparts))
        rblocks.append(np.concatenate(bparts))
    if max_len is not None:
        rows = [r[:max_len] for r in rows]
        rlabels = [r[:max_len] for r in rlabels]
        rblocks = [r[:max_len] for r in rblocks]
    L = max(len(r) for r in rows)
    p = PackPlan()
    p.L, p.t_vis = L, t_vis
    p.src = np.full((B, L), PAD_ROW, dtype=np.int32)
    p.labels = np.full((B, L), IGNORE_INDEX, dtype=np.int64)
    p.mask = np.zeros((B, L), dtype=bool)
    p.pos = np.zeros((B, L), dtype=np.int64)
    p.lengths = np.zeros(B, dtype=np.int32)
    p.kv_start = np.zeros(B, dtype=np.int32)
    p.n_blocks = n_blocks
    p.row_map = np.full(n_blocks * t_vis, -1, dtype=np.int32)
    p.vis_ids = np.full(B * L, -2, dtype=np.int32)
    for b, (r, rl, rb) in enumerate(zip(rows, rlabels, rblocks)):
        n = len(r)
        p.lengths[b] = n
        if n == 0:
            continue
        off = L - n if padding_side == "left" else 0
        p.kv_start[b] = off
        p.src[b, off:off + n] = r
        p.labels[b, off:off + n] = rl
        p.mask[b, off:off + n] = True
        p.pos[b, off:off + n] = np.arange(n)
        vis_at = np.where(rb >= 0)[0]
        table_row = rb[vis_at] * t_vis + (VISUAL_BASE - r[vis_at])
        p.row_map[table_row] = b * L + off + vis_at
        p.vis_ids[b * L + off + vis_at] = table_row
    p.desc_ids = None
    if desc_rows is not None:
        table0 = np.concatenate([[0], np.cumsum([sum(d) for d in desc_rows])]).astype(np.int64)
        flat = p.src.reshape(-1).astype(np.int64)
        sample = np.repeat(np.arange(B), L)
        p.desc_ids = np.where(flat <= DESC_BASE, table0[sample] + (DESC_BASE - flat), -2).astype(np.int32)
    return p

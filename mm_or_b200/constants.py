"""Sentinel ids and marker strings of the reference's prompt / label interface. The VALUES are part of the drop-in
contract (callers build prompts and labels with them, LLaVA/llava/constants.py:7-15); everything else here is ours."""

# labels equal to this are excluded from the loss (torch's cross-entropy ignore_index; visual and pad rows get it)
IGNORE_INDEX = -100

# placeholder ids inside input_ids that the multimodal pack replaces by embedding rows (model/pack.py):
IMAGE_TOKEN_INDEX = -200               # one per prompt: the T_vis projected visual tokens of the sample go here
VIS_DESCRIPTOR_TOKEN_INDEX = 18610     # a real vocabulary id reused as "next vis_descriptor_embs entry goes here"

# marker strings the tokenizer-side helpers of the reference look for / add (model/builder.py mirrors the additions)
DEFAULT_IMAGE_TOKEN = "<image>"
DEFAULT_IMAGE_PATCH_TOKEN = "<im_patch>"
DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN = "<im_start>", "<im_end>"

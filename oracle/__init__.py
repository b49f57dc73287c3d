"""oracle/ — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's attention semantics.

Nothing in the product path (``vats_multimodal_lm_b200``) imports this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs do, and there only as the
checker / the timed CPU baseline.

Where the algorithm lives.  The reference (S-VATS31/vats-multimodal-lm, pure Python) delegates the arithmetic of its
attention core to a third-party call that is not under /root/reference:
    torch.nn.functional.scaled_dot_product_attention   (reference pins torch==2.7.1, requirements.txt:43;
                                                        this image has torch 2.11.0)
at src/optimized_attention.py:709-714, src/transformers/vision/vit_2d/optimized_attention.py:396-402 and
src/transformers/vision/vit_3d/optimized_attention.py:302-307, and (dead code)
    flash_attn.flash_attn_interface.flash_attn_varlen_qkvpacked_func   (unpinned; flash-attn 2.8.3 docstring semantics)
at src/optimized_attention.py:628-635.  ``oracle.sdpa`` restates the published algorithm
(softmax(scale * Q K^T + mask) V with fully-masked rows -> 0) and ``oracle.mask`` restates the mask predicate.

Parity pinning.  The reference holds no golden vectors (its only numeric assertion is the causal prefix property at
tests/transformers/nlp/attention_tests.py:111-131).  The oracle is therefore pinned against OUTPUTS OF THE REFERENCE
ITSELF: ``oracle/gen_golden.py`` imports the unmodified reference modules from /root/reference in the authoring
container, runs them on seeded inputs and commits (inputs, the q/k/v entering the reference's SDPA call, its output,
the module output) under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks the oracle against every one of
them, and re-runs the reference's own prefix-consistency property.  The sliding-window and KV-cache-decode
behaviours cannot be executed by the reference at all (SURVEY.md headline facts 4-5): for those two the oracle
follows the flash-attn docstring / the intended contract and parity is "unpinned by the reference" — stated here and
in DESIGN.md.
"""
from .mask import mask_predicate  # noqa: F401
from .sdpa import sdpa_explicit, sdpa_rows, decode_explicit, expand_kv  # noqa: F401
from .prepare import decode_prepare_explicit, prefill_prepare_explicit, rope_tables  # noqa: F401

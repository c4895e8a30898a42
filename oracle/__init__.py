"""CPU oracle for the BLIP2_MR hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

A plain PyTorch fp32 restatement of the reference's algorithm for the path named by
BASELINE.json:north_star (EVA ViT-g -> ln_vision -> Q-Former -> t5_proj -> interleaved prompt ->
FlanT5-XL + LoRA loss / beam-search generate).  Every function cites the reference file:line it
follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this package; mr_blip_b200/ never does.

Pinning status
  * ViT, Q-Former, T5 (forward, loss, gradients): pinned against the reference's OWN modules
    (eva_vit.py, Qformer.py, modeling_t5.py executed unmodified through tests/golden/ref_shim.py) on
    seeded inputs; the outputs are committed under tests/golden/ by tests/golden/make_golden.py.
  * forward_mr / prompt_concatenation (blip2_mr.py:433-824): the file cannot be imported here
    (peft, tokenizer files absent); restated line by line around the pinned sub-modules.
  * LoRA (peft==0.13.0, not vendored) and beam search (transformers==4.46.1 GenerationMixin, not
    vendored; installed 5.5.0 has no .generate on PreTrainedModel): restated from the published
    algorithms -- PARITY UNPINNED for these two pieces (the reference holds no golden vectors or
    tests for any part of this path).
"""

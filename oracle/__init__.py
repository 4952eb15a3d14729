"""TEST INFRASTRUCTURE ONLY.

CPU restatement (plain PyTorch fp32) of the reference algorithm for the DPT-Hybrid-384 hot path,
plus the loader that imports the UNMODIFIED reference here for validation.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this
package; the product (omnidata_b200/) never does.

Parity status: the decoder half (modules/midas/{dpt_depth,vit,blocks}.py) is pinned against the
reference itself, imported unmodified in the build container (oracle/make_golden.py, fixtures in
tests/golden/).  The encoder half is third-party timm 0.4.12, absent from /root/reference:
its restatement is cross-checked against HuggingFace's BiT/DPT-hybrid port but is otherwise
"parity unpinned" — the reference ships no test or golden tensor for it.
"""

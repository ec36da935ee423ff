"""CPU oracle for the whisper.coreml hot path  --  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (torch fp32 / numpy / plain C), the algorithms of the
reference hot path (wangchou/whisper.coreml, `use_coreml=False` PyTorch branch).  It exists so
that the CUDA product (`whisper.coreml_b200/`) can be checked on a GPU box where
`/root/reference` does not exist.

Rules (enforced by tests/test_layout.py):
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py` (cpu_baseline / --impl reference)
    may import anything from here;
  * nothing under `whisper.coreml_b200/` imports it - the product has no CPU fallback.

Parity pinning: the oracle itself is pinned against the *reference run in the build
container* (tests/golden/make_golden.py imports /root/reference/whisper, loads the oracle's
weights through `load_state_dict`, runs the reference and stores small fixtures), and against
the reference's own known-answer tests for DTW / median filter (tests/test_timing.py of the
reference, restated in tests/test_oracle_timing.py).
"""

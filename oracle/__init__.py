"""CPU oracle for the TTA inner step.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product package
(``longcat_video_tta_b200``) never does.

Parity status: **parity unpinned by the reference** -- FifthEpoch/longcat-video-tta
ships no tests, golden vectors or known-answer fixtures for this path (SURVEY.md
section 4 / 8c) and the DiT itself (meituan-longcat/LongCat-Video, cloned un-pinned
by ``env_setup/01_setup_longcat_env.sbatch:167-170``) is not vendored.  What *is*
pinned: the loss / LoRA / optimiser-loop restatements in ``tta_oracle.py`` are
checked against the reference's own unmodified Python, imported from
``/root/reference`` through ``ref_bridge.py`` (``tests/golden/*.pt`` were generated
that way by ``make_golden.py``); the DiT arithmetic in ``dit_oracle.py`` follows
SURVEY.md Appendix A (recalled upstream behaviour) and the reference's in-tree
mirror of the upstream forward protocol (``delta_experiment/scripts/run_delta_a.py:134-217``).
"""

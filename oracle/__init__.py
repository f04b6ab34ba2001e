"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference hot path (flo-stilz/Audio-Key-Estimation:
librosa CQT call at KeyDataset.py:485-509 and PitchClassNet.forward at
models.py:651-817).  Nothing under this directory is part of the product:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` leg may import it, and only as the checker or as the
timed CPU baseline.  The product path (audio_key_estimation_b200) never
imports this package and fails loudly when its CUDA library is missing.

Parity status
-------------
* oracle.pcn_port  : PINNED.  Checked against the reference's own models.py
  (imported in the build container through oracle.ref_import) and against the
  committed golden vectors tests/golden/*.npz produced by oracle/make_golden.py.
* oracle.cqt_port  : PARITY UNPINNED.  The arithmetic lives in librosa 0.9.2 +
  resampy 0.3.1 (requirements.txt:250, :245), neither vendored in the reference
  nor installed here; the reference holds no CQT fixture.  The port restates
  the published librosa algorithm (vqt recursion) and is anchored only on
  analytic known-answer tests (pure tones, linearity).
"""

"""CPU oracle for the paired EEG/fMRI hot path -- TEST INFRASTRUCTURE, not product code.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this package, and only as the checker or the timed CPU baseline.  Nothing under
`multimodal_eeg_fmri_b200/` imports it; the product path fails loudly without its CUDA library.

What is pinned and how
----------------------
* `oracle.models` restates, functionally (plain `torch.nn.functional` calls over a reference
  `state_dict`), the reference modules on the path.  It is PINNED: `oracle/make_golden.py`
  imports the real classes from /root/reference, runs them on seeded inputs and writes
  `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks the restatement against those
  files (and, when /root/reference is present, live against the classes).
* `oracle.spectral` (window indices, band power) and `oracle.infonce` (similarity + symmetric
  InfoNCE) have NO reference implementation: **parity unpinned** for these two --
  they are authored definitions, pinned only by their own known-answer tests
  (pure tones / Parseval / scipy.signal.periodogram; orthonormal and all-equal embeddings;
  fp64 gradcheck).
"""

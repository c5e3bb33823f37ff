"""CPU oracle for the guided-source-separation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pb_chime5_b200/`` may import this
package: it is the checker used by ``tests/``, by ``__graft_entry__.smoke()``
and by the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The
product path (``pb_chime5_b200``) must fail loudly when its CUDA library is
missing and never falls back to anything in here.

Parity status of the restatement (details in DESIGN.md, section "Oracle"):

* CACGMM / mixture utils / beamformer / stable_solve: pinned against the
  reference's own code imported from ``/root/reference`` (script
  ``oracle/make_golden.py``, fixtures in ``tests/golden/``) and against the
  reference's known-answer tests (``pb_bss/tests/test_extraction/
  test_beamformer.py:182-371``).
* STFT framing / rfft layout: pinned by the doctest in
  ``pb_chime5/database/chime5/database.py:417-453``.
* WPE (``nara_wpe.wpe.wpe_v8``) and the iSTFT synthesis window: the source is
  an un-vendored third-party dependency (``nara_wpe>=0.0.6``, setup.py:142).
  The restatement follows the published algorithm; **parity unpinned** beyond
  self-consistency (perfect reconstruction for iSTFT, normal-equation residual
  for WPE).
"""

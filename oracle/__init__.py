"""CPU oracle for the asr-study acoustic hot path (TEST INFRASTRUCTURE ONLY).

This package restates, in plain numpy, the arithmetic the reference executes on
the path  preprocessing/audio.py -> core/models.py (BiLSTM) -> core/ctc_utils.py.
Every function cites the reference file:line it follows.

It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product (``asr-study_b200``) never imports it and has no CPU fallback.

PARITY PINNING STATUS
---------------------
* MFCC / log-mel front end: pinned.  ``oracle/make_golden.py`` runs the
  reference's own ``preprocessing/audio_utils.py`` (imported verbatim by path in
  the build container) and checks the restatement in ``oracle/mfcc.py`` against
  it stage by stage; the resulting vectors are committed under ``tests/golden``.
* BiLSTM / CTC / decode / Adam: **parity unpinned** by the reference itself —
  that arithmetic lives in un-vendored Keras 1.2.2 / TensorFlow 1.3.0, the
  reference ships no tests or golden vectors, and neither library is installed.
  The oracle restates the published algorithms and is cross-checked against
  independent implementations available here (brute-force CTC path enumeration,
  ``torch.nn.functional.ctc_loss``, finite differences, ``torch.optim.Adam``).
"""

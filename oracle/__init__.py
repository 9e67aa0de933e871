"""CPU oracle for the history-matching hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and there only as the checker or as the
timed CPU baseline - never on the product path (which fails loudly when the
CUDA library is missing).

Parity status
-------------
* ``oracle.analysis`` (ES / LES / IES / ILES, centring, taper, distances, prior
  sampler): PINNED.  Checked in ``tests/test_oracle_golden.py`` against golden
  vectors produced by the reference's own functions (``tests/golden/make_golden.py``
  imports ``/root/reference/notebooks/tools`` and AST-extracts the update
  functions from ``HistoryMatch.py``) and against the reference's doctest
  values and in-notebook self-checks.
* ``oracle.ressim`` (two-phase TPFA simulator): PARITY UNPINNED.  The
  simulator is the third-party dependency ``TPFA-ResSim@adc89536``
  (``/root/reference/requirements.txt:1``), which is neither vendored in the
  reference tree nor installed, and the reference holds no golden values for
  it.  The restatement follows the published Aarnes-Gimse-Lie scheme the
  notebook cites (``HistoryMatch.py:93-95``) as laid out in SURVEY.md
  Appendix A, and is pinned only by self-consistency invariants.
* ES-MDA does not exist in the reference; ``oracle.analysis.es_mda`` is this
  repo's definition on top of the reference ES update (PARITY UNPINNED except
  for the ``Na=1`` identity with ES).
"""

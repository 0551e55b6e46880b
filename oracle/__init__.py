"""CPU oracle for the CQL recommender hot path.  TEST INFRASTRUCTURE ONLY.

This package is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  ``replay_cql_b200`` never does and
fails loudly when its CUDA library is missing.

PARITY UNPINNED.  The mounted reference checkout (RePlay 0.10.0) contains
neither the CQL wrapper nor d3rlpy, d3rlpy is not pinned
in its ``pyproject.toml:31-48`` / ``poetry.lock`` and is not installable here.
The update step therefore restates the *published* d3rlpy 1.x algorithm
 and the RePlay wrapper's published behaviour
(Appendix B); the parts of the path that ARE in the checkout -- seen filtering
(``replay/models/base_rec.py:417-464``), top-k (``replay/utils.py:100-127``),
schemas (``replay/constants.py:16-31``), save/load layout
(``replay/model_handler.py:29-92``) -- are restated from those lines and pinned
by the reference's own fixtures (``tests/utils.py:59-76``) in
``tests/test_recommender_conformance.py``.
"""

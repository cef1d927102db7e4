"""TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference's hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker or
the CPU baseline -- never as the thing that is shipped or measured as the product.

Pinning status: the reference ships no tests, fixtures or golden vectors for this path (SURVEY.md
section 4), so the oracle is pinned against outputs of the reference itself: ``oracle/make_golden.py``
imports the unmodified ``/root/reference/{resnet,network}.py`` (with ``sys.modules`` stubs for the
unused ``gensim``/``clip`` imports) and scipy's ``cdist`` in the build container, and commits the
resulting vectors under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks the oracle against
them on every run.
"""

"""TEST INFRASTRUCTURE ONLY — CPU oracle for the ResLIC_TCM entropy-model hot path.

Nothing in the product package (``reslic_tcm_b200``) may import this package.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` use it, and only as the checker / the timed CPU baseline.

Parity status: the reference ships no tests and no golden vectors ("parity unpinned"
upstream).  The oracle is pinned instead against outputs of the
reference's OWN code run in the build container (``oracle/gen_golden.py`` imports
``/root/reference/src`` under stub third-party modules and extracts the TCM
``_likelihood`` twin from source); those outputs are committed under ``tests/golden``.
"""

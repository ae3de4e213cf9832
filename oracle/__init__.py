"""CPU oracle for the MOFO pretraining hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement (numpy / torch fp32)
of the reference algorithm, used as the checker by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs.  Nothing under ``mofo_b200/`` imports it; the product path is CUDA-only and
fails loudly when ``libmofo_sm100.so`` is missing.

Parity pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself, generated in the build
container by ``tests/golden/make_golden.py`` (which imports /root/reference with a
small ``timm`` shim) and committed under ``tests/golden/``.
"""

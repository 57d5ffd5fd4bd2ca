"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's LSNet training-step hot path (Duankaiwen/LSNet).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import, call, link or execute anything under this
directory — and there only as the checker / the timed CPU baseline, never as part
of the product path.  The product package ``lsnet_b200`` has no import of
``oracle`` and fails loudly when its CUDA library is missing.

Parity status: the reference's own tests hold no golden
vectors for this path, so the oracle is pinned against outputs of
the reference itself: (i) here, by importing the unmodified reference Python
(``ref_harness.py``) and committing its outputs as fixtures under
``tests/golden/`` (``tests/golden/make_golden.py``); (ii) on the GPU box, by the
reference's own CUDA DCN extension compiled into ``oracle/_ref``.
"""

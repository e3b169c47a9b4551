"""TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` exists to check the product (``goma_b200/``) and to
time the reference's CPU path next to it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import or execute it; the product path never does.
"""

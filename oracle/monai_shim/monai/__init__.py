"""Minimal stand-in for the ``monai`` package (TEST INFRASTRUCTURE).

Exposes exactly the nine names ``/root/reference/engine/utils.py:5-13`` imports so
that the reference file can be executed verbatim on CPU by
``tests/golden/make_golden.py``.  Implementations live in ``oracle/monai08.py``.
"""

"""CPU oracle for the sliding-window inference hot path of zouyunkai/MedicalSemSeg.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the CPU arm that is being *reported*, never shipped).  The
product package ``medicalsemseg_b200`` never imports this package and fails
loudly when its CUDA library is missing.

What it restates (every function cites the reference ``file:line`` it follows):

* ``oracle.monai08``        - the un-vendored MONAI 0.8.x helpers the reference
                              imports at ``engine/utils.py:5-13``
* ``oracle.sliding_window`` - ``engine/utils.py:19-159`` plus the label
                              post-processing of ``engine/test.py:140-141``
* ``oracle.vote``           - ``majority_vote.py:23-37``
* ``oracle.dice``           - MONAI ``DiceMetric`` as used at
                              ``engine/test.py:28-31,50-69``

Parity pinning status
---------------------
The reference ships NO tests, golden vectors or fixtures, so
nothing in the reference itself pins this path.  What we do instead:

* ``tests/golden/make_golden.py`` executes the reference's OWN
  ``engine/utils.py`` verbatim (imported from ``/root/reference``, never copied)
  on top of ``oracle/monai_shim`` and the reference's OWN
  ``get_class_votes``/``get_new_label`` (AST-extracted from
  ``majority_vote.py``), and stores their outputs as fixtures under
  ``tests/golden/``.  ``oracle.sliding_window`` and ``oracle.vote`` are checked
  against those fixtures bit-for-bit -> the *stitching* and *voting* arithmetic
  IS pinned to the reference's code.
* The MONAI helpers themselves (window grid, gaussian importance map, Dice)
  live in a third-party dependency that is absent from ``/root/reference`` and
  un-pinned there (``requirements.txt:1`` is the bare word ``monai``; vintage
  inferred as 0.8.x).  ``oracle.monai08`` restates their published
  algorithm from memory of MONAI 0.8.1; that part is "parity unpinned" except
  for the hand-derived known-answer tables of SURVEY.md section 8(a-1), 8(a-3).
"""

"""oracle/ -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

CPU restatement of the reference's MAML/ANIL hot path, used as the parity checker by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py``.  Nothing under ``exploring_meta_b200/`` may import this package; the product path
fails loudly when its CUDA library is missing.

PARITY UNPINNED: the reference (Kostis-S-Z/exploring_meta) ships no tests, golden vectors or
fixtures, and the arithmetic of ``clone()/adapt()`` lives in learn2learn, which is neither vendored
nor pinned (``requirements.txt`` omits it) and is not installed in this image.  The oracle is
therefore pinned as far as the material allows:

* ``oracle/ref_loader.py`` imports the reference's *own, unmodified* files
  (``core_functions/vision.py``, ``core_functions/vision_models.py``, ``utils/data_pre.py``) from
  ``/root/reference`` (build container only) on top of ``oracle/l2l_shim.py``, a restatement of the
  four learn2learn functions the path uses (published algorithm, learn2learn >= 0.1.2);
* ``oracle/maml_oracle.py`` is the self-contained functional restatement that travels to the GPU
  box; ``tests/golden/make_golden.py`` checks it against the reference-file run (fp64, <=1e-12)
  and commits the resulting vectors under ``tests/golden/``.
"""

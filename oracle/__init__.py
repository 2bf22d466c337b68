"""oracle/ -- TEST INFRASTRUCTURE ONLY (CPU restatement of the reference hot path).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``mepol_b200`` never does.
"""

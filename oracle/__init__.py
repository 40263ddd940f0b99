"""CPU oracle for the B200 GP path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product (``exauq-toolbox_b200/``)
never does.  See ``oracle/mogp_emulator/__init__.py`` for how parity is pinned.
"""

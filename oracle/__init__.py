"""CPU oracle for the separator forward pass.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product package
``puresound_b200`` never does: its ops fail loudly when the CUDA extension is
missing instead of falling back to anything in here.
"""

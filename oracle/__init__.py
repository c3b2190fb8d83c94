"""CPU oracle for the HiFi-GAN / HiFi-CAR hot path — TEST INFRASTRUCTURE ONLY.

Nothing under ``articulatory_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs use it, and there only as the checker / the timed CPU baseline.
"""

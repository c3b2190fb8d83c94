"""Entry points mirroring the reference console scripts (setup.py:52-60)."""

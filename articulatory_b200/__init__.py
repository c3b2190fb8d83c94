"""articulatory_b200 — B200-native (sm_100a) HiFi-GAN / HiFi-CAR hot path behind the
plugin surface of articulatory/articulatory.  See DESIGN.md."""
__version__ = "0.1.0"

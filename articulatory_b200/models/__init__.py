"""Plugin registry: class-name strings in YAML are resolved with
``getattr(articulatory_b200.models, config["generator_type"])`` exactly like the
reference does on ``articulatory.models`` (bin/train.py:1649-1662, utils/utils.py:325-334)."""
from .hifigan import (HiFiGANGenerator, HiFiGANMultiPeriodDiscriminator,  # noqa: F401
                      HiFiGANMultiScaleDiscriminator, HiFiGANMultiScaleMultiPeriodDiscriminator,
                      HiFiGANPeriodDiscriminator, HiFiGANScaleDiscriminator, set_default_precision)
from .inversion import BiGRU  # noqa: F401,E402

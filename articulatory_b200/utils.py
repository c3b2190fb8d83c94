"""Boundary helpers shared by the entry points (reference utils/utils.py)."""
import os

import torch
import yaml


def load_model(checkpoint, config=None, stats=None, generator2=False, precision=None):
    """Load a trained generator (reference utils/utils.py:294-372): plugin lookup by
    ``config["generator_type"]`` in ``articulatory_b200.models``, the ``upsample_kernal_sizes``
    typo work-around, ``state_dict`` from ``checkpoint["model"]["generator"]``, optional stats file
    beside the checkpoint."""
    if generator2:
        raise NotImplementedError("two-stage generator cascades are outside the B200 hot path")
    if config is None:
        with open(os.path.join(os.path.dirname(checkpoint), "config.yml")) as f:
            config = yaml.load(f, Loader=yaml.Loader)
    import articulatory_b200.models as models  # lazy, as in the reference

    name = config.get("generator_type", "ParallelWaveGANGenerator")
    if not hasattr(models, name):
        raise NotImplementedError(f"generator_type {name!r} is not on the B200 hot path "
                                  f"(available: {[n for n in dir(models) if n.startswith('HiFiGAN')]})")
    params = {k.replace("upsample_kernal_sizes", "upsample_kernel_sizes"): v
              for k, v in config["generator_params"].items()}
    if precision is not None:
        params["precision"] = precision
    model = getattr(models, name)(**params)
    model.load_state_dict(torch.load(checkpoint, map_location="cpu", weights_only=False)["model"]["generator"])
    if stats is None:
        ext = "h5" if config.get("format", "hdf5") == "hdf5" else "npy"
        cand = os.path.join(os.path.dirname(checkpoint), f"stats.{ext}")
        if os.path.exists(cand):
            stats = cand
    if stats is not None:
        model.register_stats(stats)
    if config["generator_params"].get("out_channels", 1) > 1:
        raise NotImplementedError("PQMF multi-band synthesis is outside the B200 hot path")
    return model

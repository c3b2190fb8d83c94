"""Dump-directory readers of the egs/ema/voc1 recipe (reference articulatory/datasets/audio_mel_dataset.py).

Same class names, constructor keywords and item layout as the reference, so ``run.sh --stage 2/3`` dumps feed the
B200 train / decode entry points unchanged:

* the waveform comes from the dump directory (``*.h5`` dataset ``wave`` — needs the optional ``h5py`` — or
  ``*-wave.npy``), the articulatory features from the path listed for the utterance in
  ``data/<stage>/feats.scp`` (``<stage>`` = second component of ``root_dir``, audio_mel_dataset.py:387-397);
* items are plain numpy arrays; the window cutting (``data.SpeechCollater``) and the pinned-memory hand-over to
  the fused train step happen downstream.

Speaker / phoneme side inputs (``use_spk_id``, ``use_ph``) are outside the B200 hot path.
"""
from .speech import ArtDataset, SpeechDataset, find_files, read_hdf5, read_scp  # noqa: F401

"""SpeechDataset / ArtDataset (reference datasets/audio_mel_dataset.py:305-531, 864-983)."""
import fnmatch
import logging
import os

import numpy as np


def find_files(root_dir, query="*.wav", include_root_dir=True):
    """Recursive file search (reference utils/utils.py:61-80)."""
    files = []
    for root, _, filenames in os.walk(root_dir, followlinks=True):
        for filename in fnmatch.filter(filenames, query):
            files.append(os.path.join(root, filename))
    if not include_root_dir:
        files = [f.replace(root_dir + "/", "") for f in files]
    return files


def read_hdf5(hdf5_name, hdf5_path):
    """One dataset of an hdf5 file as a numpy array (reference utils/utils.py:83-112)."""
    try:
        import h5py
    except ImportError as e:          # optional dependency, as in the reference's requirements
        raise ImportError("reading hdf5 dumps needs h5py; use `format: npy` dumps otherwise") from e
    if not os.path.exists(hdf5_name):
        raise FileNotFoundError(f"There is no such a hdf5 file ({hdf5_name}).")
    with h5py.File(hdf5_name, "r") as f:
        if hdf5_path not in f:
            raise KeyError(f"There is no such a data in hdf5 file. ({hdf5_path})")
        return f[hdf5_path][()]


def read_scp(path):
    """``utt_id value`` lines -> dict (Kaldi-style scp, as the recipe's data/<set>/feats.scp)."""
    table = {}
    with open(path) as f:
        for line in f:
            parts = line.strip().split()
            if parts:
                table[parts[0]] = parts[1] if len(parts) > 1 else ""
    return table


def _stage_of(root_dir):
    parts = root_dir.split("/")
    if len(parts) < 2:
        raise ValueError(f"root_dir {root_dir!r}: expected <dumpdir>/<stage>/... as written by the recipe")
    return parts[1]                                                     # audio_mel_dataset.py:387


def _filter_by_length(files, load_fn, threshold, what):
    lengths = [load_fn(f).shape[0] for f in files]
    idxs = [i for i in range(len(files)) if lengths[i] > threshold]
    if len(files) != len(idxs):
        logging.warning(f"Some files are filtered by {what} length threshold ({len(files)} -> {len(idxs)}).")
    return idxs


class SpeechDataset(object):
    """Audio + articulatory-feature dataset of the training recipes (audio_mel_dataset.py:305-531).

    ``__getitem__`` returns ``{'art': (T', C) float, 'audio': (T,) float[, 'utt_id': str]}`` — the dict the
    collater consumes.  ``data_dir`` (default ``"data"``, relative to the working directory like the reference)
    is where ``<stage>/feats.scp`` lives."""

    def __init__(self, root_dir, audio_query="*.h5", mel_query="*.h5", audio_load_fn=lambda x: read_hdf5(x, "wave"),
                 mel_load_fn=lambda x: read_hdf5(x, "feats"), audio_length_threshold=None, mel_length_threshold=None,
                 return_utt_id=False, allow_cache=False, transform=None, input_transform=None, output_transform=None,
                 spks=None, use_spk_id=False, use_ph=False, dataset_mode=None, data_dir="data"):
        if use_spk_id or use_ph:
            raise NotImplementedError("speaker / phoneme conditioning is outside the B200 hot path")
        if dataset_mode in ("ph2m", "m2w", "ph2a"):
            raise NotImplementedError(f"dataset_mode {dataset_mode!r} is outside the B200 hot path (a2w only)")
        audio_files = sorted(find_files(root_dir, audio_query))
        mel_files = sorted(find_files(root_dir, mel_query))
        if audio_length_threshold is not None:
            idxs = _filter_by_length(audio_files, audio_load_fn, audio_length_threshold, "audio")
            audio_files = [audio_files[i] for i in idxs]
            mel_files = [mel_files[i] for i in idxs]
        if mel_length_threshold is not None:
            idxs = _filter_by_length(mel_files, mel_load_fn, mel_length_threshold, "mel")
            audio_files = [audio_files[i] for i in idxs]
            mel_files = [mel_files[i] for i in idxs]
        assert len(audio_files) != 0, f"Not found any audio files in ${root_dir}."
        assert len(audio_files) == len(mel_files), \
            f"Number of audio and mel files are different ({len(audio_files)} vs {len(mel_files)})."
        self.audio_files, self.mel_files = audio_files, mel_files
        self.audio_load_fn, self.mel_load_fn = audio_load_fn, mel_load_fn
        if ".npy" in audio_query:
            self.utt_ids = [os.path.basename(f).replace("-wave.npy", "") for f in audio_files]
        else:
            self.utt_ids = [os.path.splitext(os.path.basename(f))[0] for f in audio_files]
        feats_path = os.path.join(data_dir, _stage_of(root_dir), "feats.scp")
        assert os.path.exists(feats_path), feats_path
        fid_to_artp = read_scp(feats_path)
        self.art_files = [fid_to_artp[fid] for fid in self.utt_ids]        # KeyError for an unlisted utterance, as the reference
        self.spks, self.spk2id, self.use_spk_id, self.use_ph = spks, None, False, False
        self.transform = transform
        self.input_transform = input_transform if input_transform is not None else transform
        self.output_transform = output_transform if output_transform is not None else transform
        self.return_utt_id = return_utt_id
        self.allow_cache = allow_cache
        self.caches = [() for _ in range(len(audio_files))] if allow_cache else None

    def __getitem__(self, idx):
        if self.allow_cache and len(self.caches[idx]) != 0:
            return self.caches[idx]
        art = np.load(self.art_files[idx])                                   # (T', C)
        if self.input_transform is not None:
            art = self.input_transform(art)
        audio = self.audio_load_fn(self.audio_files[idx])
        if self.output_transform is not None:
            audio = self.output_transform(audio)
        items = {"art": art, "audio": audio}
        if self.return_utt_id:
            items["utt_id"] = self.utt_ids[idx]
        if self.allow_cache:
            self.caches[idx] = items
        return items

    def __len__(self):
        return len(self.audio_files)


class ArtDataset(object):
    """Articulatory-feature dataset of the decoding recipes (audio_mel_dataset.py:864-983): utterance ids from the
    dumped feature files, features from ``data/<stage>/feats.scp``; ``transform == "10*f0"`` scales the pitch
    column by ten (:962-963).  Items: ``art`` or ``(utt_id, art)``."""

    def __init__(self, root_dir, mel_query="*-feats.npy", mel_length_threshold=None, mel_load_fn=np.load,
                 return_utt_id=False, allow_cache=False, transform=None, data_dir="data"):
        mel_files = sorted(find_files(root_dir, mel_query))
        if mel_length_threshold is not None:
            idxs = _filter_by_length(mel_files, mel_load_fn, mel_length_threshold, "mel")
            mel_files = [mel_files[i] for i in idxs]
        assert len(mel_files) != 0, f"Not found any mel files in ${root_dir}."
        self.mel_files, self.mel_load_fn = mel_files, mel_load_fn
        if ".npy" in mel_query:
            self.utt_ids = [os.path.basename(f).replace("-feats.npy", "") for f in mel_files]
        else:
            self.utt_ids = [os.path.splitext(os.path.basename(f))[0] for f in mel_files]
        feats_path = os.path.join(data_dir, _stage_of(root_dir), "feats.scp")
        assert os.path.exists(feats_path), feats_path
        fid_to_artp = read_scp(feats_path)
        self.art_files = [fid_to_artp[fid] for fid in self.utt_ids]
        self.transform = transform if transform is not None else ""
        self.return_utt_id = return_utt_id
        self.allow_cache = allow_cache
        self.caches = [() for _ in range(len(mel_files))] if allow_cache else None

    def __getitem__(self, idx):
        if self.allow_cache and len(self.caches[idx]) != 0:
            return self.caches[idx]
        art = np.load(self.art_files[idx])
        if self.transform == "10*f0":
            art[:, 0] *= 10
        items = (self.utt_ids[idx], art) if self.return_utt_id else art
        if self.allow_cache:
            self.caches[idx] = items
        return items

    def __len__(self):
        return len(self.mel_files)

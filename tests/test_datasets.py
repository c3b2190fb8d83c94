"""CPU tests of the dump-directory readers (reference datasets/audio_mel_dataset.py:305-531, 864-983): the
recipe's layout — dump/<stage>/norm/*-wave.npy (+ *-feats.npy) and data/<stage>/feats.scp — built in a tmp dir."""
import os

import numpy as np
import pytest


def _make_recipe(tmp_path, n=3, stage="train_nodev"):
    rng = np.random.default_rng(0)
    dump = tmp_path / "dump" / stage / "norm"
    data = tmp_path / "data" / stage
    raw = tmp_path / "raw_art"
    for d in (dump, data, raw):
        d.mkdir(parents=True)
    lines, utts = [], []
    for i in range(n):
        utt = f"utt{i:02d}"
        frames = 50 + 10 * i
        np.save(dump / f"{utt}-wave.npy", rng.standard_normal(frames * 80).astype(np.float32))
        np.save(dump / f"{utt}-feats.npy", rng.standard_normal((frames, 80)).astype(np.float32))   # mel, unused for a2w
        art = rng.standard_normal((frames, 13)).astype(np.float32)
        np.save(raw / f"{utt}.npy", art)
        lines.append(f"{utt} {raw / (utt + '.npy')}")
        utts.append((utt, art))
    (data / "feats.scp").write_text("\n".join(reversed(lines)) + "\n")      # scp order != file order
    return dump, utts


def test_speech_dataset_reads_wave_from_dump_and_art_from_scp(tmp_path, monkeypatch):
    from articulatory_b200.datasets import SpeechDataset
    dump, utts = _make_recipe(tmp_path)
    monkeypatch.chdir(tmp_path)
    ds = SpeechDataset(os.path.join("dump", "train_nodev", "norm"), audio_query="*-wave.npy", mel_query="*-feats.npy",
                       audio_load_fn=np.load, mel_load_fn=np.load, return_utt_id=True, dataset_mode="a2w")
    assert len(ds) == 3 and ds.utt_ids == [u for u, _ in utts]
    for i, (utt, art) in enumerate(utts):
        it = ds[i]
        assert it["utt_id"] == utt
        np.testing.assert_array_equal(it["art"], art)
        np.testing.assert_array_equal(it["audio"], np.load(dump / f"{utt}-wave.npy"))
        assert len(it["audio"]) == 80 * len(it["art"])


def test_speech_dataset_length_threshold_transform_and_cache(tmp_path, monkeypatch):
    from articulatory_b200.datasets import SpeechDataset
    _make_recipe(tmp_path)
    monkeypatch.chdir(tmp_path)
    ds = SpeechDataset("dump/train_nodev/norm", audio_query="*-wave.npy", mel_query="*-feats.npy", audio_load_fn=np.load,
                       mel_load_fn=np.load, audio_length_threshold=50 * 80, allow_cache=True,
                       input_transform=lambda a: a * 2.0)
    assert len(ds) == 2 and ds.utt_ids == ["utt01", "utt02"]            # the 50-frame utterance is filtered (">")
    a = ds[0]
    assert ds[0] is a                                                    # cached item
    np.testing.assert_allclose(a["art"], 2.0 * np.load(ds.art_files[0]))


def test_speech_dataset_errors(tmp_path, monkeypatch):
    from articulatory_b200.datasets import SpeechDataset
    _make_recipe(tmp_path)
    monkeypatch.chdir(tmp_path)
    kw = dict(audio_query="*-wave.npy", mel_query="*-feats.npy", audio_load_fn=np.load, mel_load_fn=np.load)
    with pytest.raises(NotImplementedError):
        SpeechDataset("dump/train_nodev/norm", use_spk_id=True, **kw)
    with pytest.raises(AssertionError):
        SpeechDataset("dump/train_nodev/empty", **kw)                     # no audio files
    os.remove("data/train_nodev/feats.scp")
    with pytest.raises(AssertionError):
        SpeechDataset("dump/train_nodev/norm", **kw)                      # reference asserts the scp exists


def test_art_dataset_and_f0_transform(tmp_path, monkeypatch):
    from articulatory_b200.datasets import ArtDataset
    _, utts = _make_recipe(tmp_path, stage="eval")
    monkeypatch.chdir(tmp_path)
    ds = ArtDataset("dump/eval/norm", return_utt_id=True)
    assert len(ds) == 3
    for i, (utt, art) in enumerate(utts):
        u, a = ds[i]
        assert u == utt
        np.testing.assert_array_equal(a, art)
    ds10 = ArtDataset("dump/eval/norm", transform="10*f0")
    a = ds10[1]
    np.testing.assert_allclose(a[:, 0], 10 * utts[1][1][:, 0], rtol=1e-6)
    np.testing.assert_array_equal(a[:, 1:], utts[1][1][:, 1:])


def test_train_entry_point_item_loader(tmp_path, monkeypatch):
    """bin/train.py::_load_items: the recipe layout goes through SpeechDataset (features from the scp), a bare dump
    directory falls back to the in-dump wave / feats pairs."""
    from articulatory_b200.bin.train import _load_items
    dump, utts = _make_recipe(tmp_path)
    monkeypatch.chdir(tmp_path)
    items = _load_items("dump/train_nodev/norm", {"format": "npy", "dataset_mode": "a2w"})
    assert len(items) == 3 and items[0]["art"].shape[1] == 13
    np.testing.assert_array_equal(items[2]["art"], utts[2][1])
    os.remove("data/train_nodev/feats.scp")
    items = _load_items("dump/train_nodev/norm", {"format": "npy"})
    assert len(items) == 3 and items[0]["art"].shape[1] == 80               # the dump's own feats
    with pytest.raises(ValueError):
        _load_items("dump/train_nodev/norm", {"format": "wav"})

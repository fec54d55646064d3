"""The `--mode train/generate` command line (new_scripy.py:1292-1321): flags, defaults, error behaviour on
CPU; a tiny train -> checkpoint -> generate round trip on the GPU."""
import os

import pytest
import torch

from diffusionmodel_b200 import cli


def test_flags_and_defaults_match_the_reference(capsys):
    with pytest.raises(SystemExit) as e:
        cli.main(["--mode", "generate"])                    # new_scripy.py:1312-1315: help + exit(1)
    assert e.value.code == 1
    assert "Checkpoint path required" in capsys.readouterr().out
    with pytest.raises(SystemExit):
        cli.main(["--mode", "bogus"])
    assert cli.Cfg.GUIDE_SCALES == [2.0, 4.0] and cli.Cfg.SAMPLES_PER_CLASS == 3
    assert (cli.Cfg.BATCH_SIZE, cli.Cfg.ACCUM_STEPS, cli.Cfg.N_T, cli.Cfg.N_FEAT) == (4, 4, 700, 192)


def test_synthetic_batch_matches_dataset_contract():
    g = torch.Generator().manual_seed(0)
    x, c, m = cli.synth_batch(g, 4, 64, 5)
    assert x.shape == (4, 3, 64, 64) and float(x.min()) >= -1 and float(x.max()) <= 1
    assert c.dtype == torch.int64 and int(c.max()) < 5
    assert set(m.unique().tolist()) <= {0.5, 1.0, 3.0} and (m == 3.0).any()      # new_scripy.py:535-546


@pytest.mark.gpu
def test_train_then_generate_roundtrip(dev, tmp_path, monkeypatch):
    monkeypatch.setattr(cli.Cfg, "SAVE_DIR", str(tmp_path / "ckpt") + "/")
    monkeypatch.setattr(cli.Cfg, "SAMPLE_DIR", str(tmp_path / "samples") + "/")
    monkeypatch.setattr(cli.Cfg, "N_T", 20)
    cli.main(["--mode", "train", "--epochs", "1", "--steps_per_epoch", "8", "--n_feat", "16", "--img", "128"])
    ckpt = os.path.join(cli.Cfg.SAVE_DIR, "best_model.pt")
    sd = torch.load(ckpt)["model_state_dict"]
    assert len(sd) == 415
    out = cli.gen_samples(ckpt, n_samples_per_class=1, guide_scales=[2.0], n_classes=5, n_feat=16, img=128)
    assert out[2.0].shape == (5, 3, 128, 128) and torch.isfinite(out[2.0]).all()
    assert cli.gen_samples(str(tmp_path / "missing.pt")) is None             # new_scripy.py:967-969

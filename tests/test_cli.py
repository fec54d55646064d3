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


def _write_voc_dir(root, classes=("D00", "D10", "D20", "D40", "Repair"), per_class=2, seed=3):
    """A throw-away dataset in the reference's layout: images/<class>/*.png + annotations/*.xml (new_scripy.py:496-511)."""
    import numpy as np
    from PIL import Image
    rng = np.random.RandomState(seed)
    os.makedirs(os.path.join(root, "annotations"), exist_ok=True)
    for cname in classes:
        os.makedirs(os.path.join(root, "images", cname), exist_ok=True)
        for k in range(per_class):
            w, h = int(rng.randint(90, 200)), int(rng.randint(90, 200))
            img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
            stem = f"{cname}_{k}"
            Image.fromarray(img).save(os.path.join(root, "images", cname, stem + ".png"))
            x0, y0 = int(rng.randint(0, w // 2)), int(rng.randint(0, h // 2))
            with open(os.path.join(root, "annotations", stem + ".xml"), "w") as f:
                f.write(f"<annotation><size><width>{w}</width><height>{h}</height></size><object><bndbox><xmin>{x0}</xmin>"
                        f"<ymin>{y0}</ymin><xmax>{x0 + w // 3}</xmax><ymax>{y0 + h // 3}</ymax></bndbox></object></annotation>")


@pytest.mark.gpu
def test_train_and_generate_from_a_dataset_directory_with_quality_metrics(dev, tmp_path, monkeypatch):
    """--data DIR: training batches from the cached dataset (n_classes = its class folders, new_scripy.py:692), and
    generate's quality assessment against real images (SSIM / PSNR; FID is NaN without a feature network and needs >= 10
    samples) written to quality_metrics.json like new_scripy.py:1085-1101."""
    import json
    import math
    monkeypatch.setattr(cli.Cfg, "SAVE_DIR", str(tmp_path / "ckpt") + "/")
    monkeypatch.setattr(cli.Cfg, "SAMPLE_DIR", str(tmp_path / "samples") + "/")
    monkeypatch.setattr(cli.Cfg, "N_T", 12)
    data = str(tmp_path / "cropped_images")
    _write_voc_dir(data)
    cli.main(["--mode", "train", "--epochs", "1", "--steps_per_epoch", "4", "--n_feat", "16", "--img", "128", "--data", data])
    ckpt = os.path.join(cli.Cfg.SAVE_DIR, "best_model.pt")
    out = cli.gen_samples(ckpt, n_samples_per_class=2, guide_scales=[2.0, 4.0], n_classes=5, n_feat=16, img=128, data=data)
    assert out[2.0].shape == (10, 3, 128, 128) and out[4.0].shape == (10, 3, 128, 128)
    qm = out["quality_metrics"]
    assert set(qm) == {2.0, 4.0}
    for m in qm.values():                       # 2 * min(5, 4) = 8 real images: fewer than 10, so no FID key (:1266)
        assert set(m) == {"ssim", "psnr"} and all(math.isfinite(v) for v in m.values())
    saved = json.load(open(os.path.join(cli.Cfg.SAMPLE_DIR, "quality_metrics.json")))
    assert set(saved) == {"2.0", "4.0"} and saved["2.0"]["psnr"] == pytest.approx(qm[2.0]["psnr"])
    out = cli.gen_samples(ckpt, n_samples_per_class=1, guide_scales=[2.0], n_classes=5, n_feat=16, img=128)      # no --data
    assert out["quality_metrics"] == {}

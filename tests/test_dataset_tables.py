"""CPU: the reference's label tables as 256-entry LUTs (dataset.py:10-63, utils.py:155-173), and
GPU: the dataset-preparation caller of the sampling loop (utils.py:176-208)."""
import os

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import _refshim


def _loop_remap(arr, mapping, dtype=np.uint8):
    """The reference's remap, literally (dataset.py:53-61 / utils.py:170-172)."""
    out = np.zeros_like(arr, dtype=dtype)
    for k, v in mapping.items():
        out[arr == k] = v
    return out


def test_lut_equals_the_reference_loop(tmp_path):
    from ldiffusion_b200 import dataset as D
    rng = np.random.default_rng(0)
    arr = rng.integers(0, 256, (37, 53)).astype(np.uint8)
    keys = sorted(set(D.pixel_to_label) | set(D.pixel_to_label_cell) | set(D.ID_TO_CLASS))
    arr.flat[:len(keys)] = keys                            # every table key occurs at least once
    for table in (D.pixel_to_label, D.pixel_to_label_cell, D.ID_TO_CLASS):
        assert np.array_equal(D.label_lut_numpy(table)[arr], _loop_remap(arr, table))
    assert np.array_equal(D.map_mask(arr), _loop_remap(arr, D.ID_TO_CLASS, np.int64)) and D.map_mask(arr).dtype == np.int64
    p = tmp_path / "lab.png"
    Image.fromarray(arr).save(p)
    assert np.array_equal(D.convert_labels(str(p), "tissue"), _loop_remap(arr, D.pixel_to_label))
    assert np.array_equal(D.convert_labels(str(p), "cell"), _loop_remap(arr, D.pixel_to_label_cell))
    with pytest.raises(ValueError, match="Unsupported level"):
        D.convert_labels(str(p), "organ")
    with pytest.raises(ValueError):
        D.label_lut_numpy({300: 1})


@pytest.mark.skipif(not _refshim.available(), reason="reference checkout not present (GPU box)")
def test_tables_are_the_references():
    from ldiffusion_b200 import dataset as D
    src = open(os.path.join(_refshim.REF_ROOT, "dataset.py")).read()
    ns = {}
    for name in ("pixel_to_label", "pixel_to_label_cell", "ID_TO_CLASS"):       # the three dict literals only:
        start = src.index(f"{name} = {{")                                       # dataset.py imports cv2 / torchvision
        exec(src[start:src.index("}", start) + 1], ns)
        assert getattr(D, name) == ns[name]


def test_convert_and_save_label_and_plain_copy(tmp_path):
    from ldiffusion_b200 import dataset as D, utils as U
    rng = np.random.default_rng(1)
    arr = rng.choice(np.array(list(D.pixel_to_label), dtype=np.uint8), (40, 24))
    U.convert_and_save_label(Image.fromarray(arr), tmp_path / "l.png", D.pixel_to_label)
    assert np.array_equal(np.array(Image.open(tmp_path / "l.png")), _loop_remap(arr, D.pixel_to_label))
    arr32 = arr.astype(np.int32)                                                # 'I' mode image: the literal loop
    U.convert_and_save_label(Image.fromarray(arr32), tmp_path / "l32.png", D.pixel_to_label)
    assert np.array_equal(np.array(Image.open(tmp_path / "l32.png")), _loop_remap(arr, D.pixel_to_label))
    for i, size in enumerate([(30, 20), (30, 20), (31, 20)]):
        Image.fromarray(np.zeros((size[1], size[0], 3), np.uint8)).save(tmp_path / f"i{i}.png")
    assert U.check_images_same_size([tmp_path / "i0.png", tmp_path / "i1.png"])
    assert not U.check_images_same_size([tmp_path / f"i{i}.png" for i in range(3)])
    U.copy_or_convert_image(None, tmp_path / "i0.png", tmp_path / "copy.png", use_diffusion=False)
    assert open(tmp_path / "copy.png", "rb").read() == open(tmp_path / "i0.png", "rb").read()


@pytest.mark.gpu
def test_gt_lut_from_the_level_tables_matches_remapped_labels():
    """confusion matrix of raw PUMA gray levels through the fused table == matrix of the remapped labels."""
    import ldiffusion_b200 as L
    from ldiffusion_b200 import dataset as D
    from oracle import metrics as omet
    rng = np.random.default_rng(2)
    for table, K in ((D.pixel_to_label, 7), (D.pixel_to_label_cell, 11)):
        raw = rng.choice(np.array(list(table) + [33], dtype=np.uint8), (2, 96, 80))      # 33: not in the table -> 0
        pred = rng.integers(0, K, raw.shape).astype(np.uint8)
        C = L.confusion_matrix(torch.from_numpy(pred).cuda(), torch.from_numpy(raw).cuda(), K,
                               gt_lut=D.label_lut(table))
        assert np.array_equal(C.cpu().numpy(), omet.confusion_matrix(pred, _loop_remap(raw, table), K))


@pytest.mark.gpu
def test_copy_or_convert_image_runs_the_one_step_loop(tmp_path):
    """utils.py:176-208 with the stand-in pipeline: the saved PNG is the decode tail of the captured tensor."""
    from ldiffusion_b200 import utils as U
    from ldiffusion_b200.standin import StandInPipeline
    from oracle import decode_tail as odt
    pipe = StandInPipeline("cuda", seed=7)
    dec = []
    vae_dec = pipe.vae.decode

    def decode(z):
        out = vae_dec(z); dec.append(out.sample.detach().cpu()); return out

    pipe.vae.decode = decode
    rng = np.random.default_rng(3)
    img = Image.fromarray(rng.integers(0, 256, (200, 200, 3), dtype=np.uint8))
    U.copy_or_convert_image(img, "unused", tmp_path / "out.png", pipe, pipe.unet, use_diffusion=True)
    saved = np.array(Image.open(tmp_path / "out.png"))
    assert saved.shape == (1024, 1024, 3) and len(dec) == 1
    assert np.array_equal(saved, odt.decode_tail_chain(dec[0])[0])

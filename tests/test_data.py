"""CPU tests of the host-side loader: image -> matrix adapter (rcn/src/lib.rs:27-41), error type
(rcn/src/errors.rs) and the class-directory sampling rules of load_data (rcn/src/rcn.rs:367-404). PNG decoding itself is
Pillow's job on the host; the GPU receives the u8 buffers these functions return."""
import numpy as np
import pytest
from PIL import Image

from mercer_research_b200 import data as D


def test_get_pixel_matrix_luma_and_luma_alpha():
    a = np.arange(12, dtype=np.uint8).reshape(3, 4)
    m = D.get_pixel_matrix(Image.fromarray(a, mode="L"))
    assert m.dtype == np.float64 and m.shape == (3, 4) and np.array_equal(m, a)      # (r, c) = pixel (x=c, y=r), 0..255 as f64
    la = np.stack([a, 255 - a], axis=-1)
    assert np.array_equal(D.get_pixel_matrix(Image.fromarray(la, mode="LA")), a)     # alpha ignored (lib.rs:34-38)


def test_non_grayscale_is_the_reference_error():
    rgb = Image.fromarray(np.zeros((2, 2, 3), dtype=np.uint8), mode="RGB")
    with pytest.raises(D.InvalidGrayscaleImageError) as e:
        D.get_pixel_matrix(rgb)
    assert "not Luma8" in str(e.value)                                                # errors.rs Display text


@pytest.fixture()
def tree(tmp_path):
    rng = np.random.default_rng(0)
    pixels = {}
    for cls in ("2", "0", "1"):                      # created out of order: load_data sorts the class directories
        d = tmp_path / "set" / cls
        d.mkdir(parents=True)
        for k in range(5):
            a = rng.integers(0, 256, size=(6, 7), dtype=np.uint8)
            a[0, 0] = int(cls) * 50 + k              # tag: class and file index
            mode = "L" if k % 2 == 0 else "RGB"      # colour files are converted by .grayscale() (rcn.rs:398)
            img = Image.fromarray(a, mode="L").convert(mode)
            img.save(d / f"{k}.png")
            pixels[(int(cls), k)] = a
    return str(tmp_path / "set"), pixels


def test_load_data_sampling_rules(tree):
    path, pixels = tree
    images, labels = D.load_data(path, 3, np.random.default_rng(1))
    assert images.shape == (9, 6, 7) and images.dtype == np.uint8
    assert labels.tolist() == [0, 0, 0, 1, 1, 1, 2, 2, 2]                             # class index = sorted directory order
    seen = set()
    for img, lab in zip(images, labels):
        cls, k = divmod(int(img[0, 0]), 50)
        assert cls == lab and np.array_equal(img, pixels[(cls, k)])                    # grey RGB files decode to the same bytes
        seen.add((cls, k))
    assert len(seen) == 9                                                              # drawn WITHOUT replacement (paths.remove)
    again, _ = D.load_data(path, 3, np.random.default_rng(1))
    assert np.array_equal(images, again)                                               # the harness owns the seed
    full, _ = D.load_data(path, 5, np.random.default_rng(2))
    assert len({int(i[0, 0]) for i in full}) == 15                                     # limit == class size: every file once


def test_load_data_limit_too_large_is_the_reference_panic(tree):
    path, _ = tree
    with pytest.raises(ValueError) as e:
        D.load_data(path, 6, np.random.default_rng(0))
    assert str(e.value) == f"provided class_size_limit for {path} too large! expected 6 <= 5"   # rcn.rs:383-390


def test_grayscale_uses_the_image_crates_integer_luma_not_rec601(tmp_path):
    """`.grayscale()` of the `image` crate (rcn.rs:83,398): (2126 R + 7152 G + 722 B) / 10000, truncating -- a colour PNG
    must give the reference's pixels, not Pillow's Rec.601 convert("L"); 16-bit input is the reference's error."""
    rng = np.random.default_rng(3)
    rgb = rng.integers(0, 256, size=(5, 6, 3), dtype=np.uint8)
    r32 = rgb.astype(np.uint32)
    want = ((2126 * r32[..., 0] + 7152 * r32[..., 1] + 722 * r32[..., 2]) // 10000).astype(np.uint8)
    Image.fromarray(rgb, mode="RGB").save(tmp_path / "c.png")
    got = D.load_grayscale(str(tmp_path / "c.png"))
    assert np.array_equal(got, want)
    assert not np.array_equal(got, np.asarray(Image.fromarray(rgb, mode="RGB").convert("L")))   # Rec.601 differs
    rgba = np.concatenate([rgb, rng.integers(0, 256, size=(5, 6, 1), dtype=np.uint8)], axis=-1)
    Image.fromarray(rgba, mode="RGBA").save(tmp_path / "a.png")
    assert np.array_equal(D.load_grayscale(str(tmp_path / "a.png")), want)                      # alpha ignored (lib.rs:34-38)
    pal = Image.fromarray(rgb, mode="RGB").quantize(8)
    pal.save(tmp_path / "p.png")
    p_rgb = np.asarray(pal.convert("RGB"), dtype=np.uint32)
    assert np.array_equal(D.load_grayscale(str(tmp_path / "p.png")),
                          ((2126 * p_rgb[..., 0] + 7152 * p_rgb[..., 1] + 722 * p_rgb[..., 2]) // 10000).astype(np.uint8))
    Image.fromarray(rng.integers(0, 65536, size=(4, 4), dtype=np.uint16)).save(tmp_path / "s.png")   # 16-bit grey
    with pytest.raises(D.InvalidGrayscaleImageError):
        D.load_grayscale(str(tmp_path / "s.png"))

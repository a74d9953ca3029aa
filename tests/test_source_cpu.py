"""CPU suite: the frame source (System::NextFrame_kitti replacement) decodes PNG (every row filter, written by a real encoder)
and PGM into batches in sequence order and ends the run at the first missing pair.  No GPU needed: without a device the ring is
ordinary memory."""
import os

import numpy as np
import pytest


def _seq(pkg, n, w, h):
    return pkg.synth_sequence(11, 3, n, w, h)


@pytest.mark.parametrize("fmt", ["png", "pgm"])
def test_source_roundtrip_own_writers(pkg, tmp_path, fmt):
    w, h, n, B = 333, 97, 11, 4
    left, right = _seq(pkg, n, w, h)
    pkg.write_sequence(str(tmp_path), left, right, first_index=5, fmt=fmt)
    src = pkg.FrameSource(str(tmp_path), w, h, B, ring_batches=3, threads=3, first_index=5)
    assert src.frame_count() == n
    got_l, got_r = [], []
    while True:
        b = src.next()
        if b is None:
            break
        assert 1 <= len(b[0]) <= B
        got_l.append(b[0].copy())
        got_r.append(b[1].copy())
    assert [len(x) for x in got_l] == [4, 4, 3]
    assert np.array_equal(np.concatenate(got_l), left) and np.array_equal(np.concatenate(got_r), right)
    assert src.next() is None  # stays at the end
    src.close()


def test_source_reads_png_of_a_real_encoder(pkg, tmp_path):
    cv2 = pytest.importorskip("cv2")
    w, h, n = 640, 200, 6
    left, right = _seq(pkg, n, w, h)
    rng = np.random.default_rng(0)
    left[2] = rng.integers(0, 256, (h, w), dtype=np.uint8)  # noise: the encoder picks other row filters than on smooth images
    for side, imgs in ((0, left), (1, right)):
        os.makedirs(tmp_path / f"image_{side}")
        for i in range(n):
            assert cv2.imwrite(str(tmp_path / f"image_{side}" / f"{i:06d}.png"), imgs[i], [cv2.IMWRITE_PNG_COMPRESSION, 3 + i % 5,
                                                                                           cv2.IMWRITE_PNG_STRATEGY, i % 4])
    os.remove(tmp_path / "image_1" / "000005.png")  # the run ends at the first missing pair (System.cpp:88-93)
    src = pkg.FrameSource(str(tmp_path), w, h, 2, ring_batches=3, threads=4)
    assert src.frame_count() == 5
    out_l, out_r = [], []
    while (b := src.next()) is not None:
        out_l.append(b[0].copy())
        out_r.append(b[1].copy())
    assert np.array_equal(np.concatenate(out_l), left[:5]) and np.array_equal(np.concatenate(out_r), right[:5])
    # cv2 reads our own PNG writer's files back identically too
    pkg.write_sequence(str(tmp_path / "own"), left[:2], right[:2], fmt="png", zlib_level=6)
    back = cv2.imread(str(tmp_path / "own" / "image_0" / "000001.png"), cv2.IMREAD_GRAYSCALE)
    assert np.array_equal(back, left[1])
    src.close()


def test_source_errors_are_loud(pkg, tmp_path):
    cv2 = pytest.importorskip("cv2")
    w, h = 64, 48
    left, right = _seq(pkg, 3, w, h)
    pkg.write_sequence(str(tmp_path), left, right)
    cv2.imwrite(str(tmp_path / "image_0" / "000001.png"), np.zeros((h, w, 3), np.uint8))  # a colour PNG is not silently converted
    src = pkg.FrameSource(str(tmp_path), w, h, 2, ring_batches=3, threads=2)
    with pytest.raises(pkg.SvoError) as e:
        while src.next() is not None:
            pass
    assert "greyscale" in str(e.value)
    src.close()
    src = pkg.FrameSource(str(tmp_path), w + 1, h, 2, ring_batches=3, threads=2)  # wrong size
    with pytest.raises(pkg.SvoError):
        src.next()
    src.close()
    with pytest.raises(pkg.SvoError):
        pkg.FrameSource(str(tmp_path), w, h, 2, ring_batches=2, threads=2)  # ring needs >= 3 batches (two stay lent)


def test_source_corrupt_png_never_crashes(pkg, tmp_path):
    """Flipped bytes, truncations, overwritten chunk headers and trailing garbage in a PNG end in SvoError or in a clean decode --
    never in a crash or a hang (the decoder threads read files the reference would hand to cv::imread, System.cpp:79-86)."""
    w, h = 97, 61
    left, right = _seq(pkg, 2, w, h)
    pkg.write_sequence(str(tmp_path), left, right, fmt="png", zlib_level=6)
    path = tmp_path / "image_0" / "000000.png"
    good = path.read_bytes()
    rng = np.random.default_rng(5)
    decoded = failed = 0
    for it in range(120):
        b = bytearray(good)
        mode = it % 4
        if mode == 0:
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(8, len(b)))] = int(rng.integers(0, 256))
        elif mode == 1:
            b = b[:int(rng.integers(0, len(b)))]
        elif mode == 2:
            p = int(rng.integers(8, len(b) - 8))
            b[p:p + 4] = int(rng.integers(0, 2**32)).to_bytes(4, "big")
        else:
            b += bytes(rng.integers(0, 256, int(rng.integers(1, 64)), dtype=np.uint8))
        path.write_bytes(bytes(b))
        try:
            src = pkg.FrameSource(str(tmp_path), w, h, 2, ring_batches=3, threads=2)
        except pkg.SvoError:
            failed += 1
            continue
        try:
            got = src.next()
            if got is not None and mode == 3:  # trailing bytes after IEND do not change the image
                assert np.array_equal(got[0][0], left[0])
            decoded += 1
        except pkg.SvoError:
            failed += 1
        src.close()
    assert decoded >= 30 and failed >= 60  # every appended-garbage case decodes, (nearly) every damaged stream is refused

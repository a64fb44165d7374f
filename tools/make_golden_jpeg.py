#!/usr/bin/env python
"""Golden vectors for the JPEG leg of tw_decode_gray: small JPEG files written by cv2.imencode in every
coding variant the decoder supports, with the gray image cv2.imdecode(..., IMREAD_GRAYSCALE) returns for each (cv2 4.13.0,
libjpeg-turbo).  Run in the build container (needs cv2); writes tests/golden/jpg/*.jpg + *.gray.npy.
tests/golden/jpg/fixture_s1_capture1.jpg is the reference's own test file (test/fixture/*/scenario1/capture1.jpg, the three
revisions are byte-identical); its expected pixels are tests/golden/fixture_s1_expected.npy."""
import os
import cv2
import numpy as np

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "jpg")
SS = {"444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420,
      "440": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440, "411": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411}


def picture(h, w, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    a = np.stack([(x * 3 + y) % 256, (x + y * 2) % 256, (x * y // 7) % 256], -1).astype(np.uint8)
    a[h // 4:h // 2, w // 4:w // 2] = (255, 0, 0)
    a[::7, ::5] = rng.integers(0, 256, a[::7, ::5].shape, dtype=np.uint8)
    return a


# name: (h, w, quality, progressive, sampling, restart interval, optimize, gray source)
CASES = {
    "base_420_q75": (61, 83, 75, 0, "420", 0, 0, False),
    "base_444_q95_opt": (40, 40, 95, 0, "444", 0, 1, False),
    "base_422_rst3": (37, 53, 60, 0, "422", 3, 0, False),
    "base_411_q30": (48, 70, 30, 0, "411", 0, 0, False),
    "base_440_q100": (33, 17, 100, 0, "440", 0, 0, False),
    "prog_420_q75": (61, 83, 75, 1, "420", 0, 0, False),
    "prog_444_q90_rst2": (29, 45, 90, 1, "444", 2, 0, False),
    "prog_422_q40": (64, 64, 40, 1, "422", 0, 1, False),
    "gray_base_q80": (50, 31, 80, 0, None, 0, 0, True),
    "gray_prog_q80_rst4": (50, 31, 80, 1, None, 4, 0, True),
    "tiny_1x1": (1, 1, 75, 0, "420", 0, 0, False),
}

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for i, (name, (h, w, q, prog, ss, rst, opt, gray)) in enumerate(CASES.items()):
        a = picture(h, w, i)
        if gray:
            a = cv2.cvtColor(a, cv2.COLOR_BGR2GRAY)
        params = [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_PROGRESSIVE, prog, cv2.IMWRITE_JPEG_OPTIMIZE, opt, cv2.IMWRITE_JPEG_RST_INTERVAL, rst]
        if ss:
            params += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SS[ss]]
        ok, enc = cv2.imencode(".jpg", a, params)
        assert ok
        open(os.path.join(OUT, name + ".jpg"), "wb").write(enc.tobytes())
        np.save(os.path.join(OUT, name + ".gray.npy"), cv2.imdecode(enc, cv2.IMREAD_GRAYSCALE))
        print(name, len(enc))

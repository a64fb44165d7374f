"""CPU test (no GPU) of the JS-level mirror's host logic (tidal-wave_b200/tidalwave.py, /root/reference/index.js:14-73,
src/broker.cpp:29-86): request order of the events, the reference's error messages and the Report, with the dispatcher replaced by
a stub that decodes the two paths the way the C++ decoder threads do (tw_decode_gray; the real dispatcher, decoder threads
included, is covered by tests/test_gpu_tidalwave_api.py and tests/test_gpu_parity.py::test_pool_path_requests)."""
import os

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class StubPool:
    def __init__(self, tw):
        self.tw, self.reqs, self.answers = tw, [], []

    def request_files(self, expected, target):  # tw_pool_submit_files: imread both, "Can't open <path>" (src/opticalflow.cpp:37-49)
        a = self.tw.imread_gray(expected)
        b = self.tw.imread_gray(target) if a is not None else None
        if a is None or b is None:
            self.answers.append({"status": "ERROR", "reason": "Can't open " + (expected if a is None else target), "vector": []})
        else:
            self.reqs.append((a.shape, b.shape))
            h, w = a.shape
            self.answers.append({"status": "OK", "n_vectors": 0, "vector": [], "width": w, "height": h, "span": 10, "threshold": 5.0, "time": 0.0})
        return len(self.answers) - 1

    def wait(self, rid):
        return self.answers[rid]

    def stop(self):
        pass

    def close(self):
        pass


def test_decode_workers_keep_request_order_and_messages(tw, tmp_path):
    t = tw.TidalWave({"numThreads": 3})
    stub = StubPool(tw)
    t._ensure_pool = lambda w, h: (setattr(t, "_pool", stub), stub)[1]
    (tmp_path / "junk.jpg").write_bytes(b"\xff\xd8\xff\xe0junk")
    s1 = os.path.join(GOLD, "jpg", "fixture_s1_capture1.jpg")
    s2e, s2r = os.path.join(GOLD, "png", "fixture_s2_expected.png"), os.path.join(GOLD, "png", "fixture_s2_revision2.png")
    ev = []
    t.on("data", lambda r: ev.append(("data", r["expect_image"], r["target_image"], r["width"], r["height"])))
    t.on("error", lambda r: ev.append(("error", r["reason"]))).on("finish", lambda r: ev.append(("finish", r)))
    t.calc(s1, s1)
    t.calc("", "x")
    t.calc(s2e, "")
    t.calc(s2e, str(tmp_path / "junk.jpg"))
    t.calc(s2e, s2r)
    t.calc(str(tmp_path / "missing.png"), s2r)
    t.dispose()
    t.calc(s1, s1)  # after dispose: ignored
    assert ev == [("data", s1, s1, 280, 279), ("error", "ExpectImagePath is empty."), ("error", "TargetImagePath is empty."),
                  ("error", "Can't open " + str(tmp_path / "junk.jpg")), ("data", s2e, s2r, 180, 117),
                  ("error", "Can't open " + str(tmp_path / "missing.png")), ("finish", {"request": 6, "data": 2, "error": 4})]
    assert stub.reqs == [((279, 280), (279, 280)), ((117, 180), (117, 180))]

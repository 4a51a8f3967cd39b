"""Host logic of batch_filter and of the plan cache, without a GPU: process_img / the native Plan are replaced by markers so
that only scheduling, failure semantics (reference pystripe/core.py:1687-1771) and cache policy are exercised."""
import threading
import time
from pathlib import Path

import numpy as np
import pytest

from pystripe import _io, _native, core


@pytest.fixture
def host_only(monkeypatch):
    monkeypatch.setattr(core, "_batch_buffer", lambda device, shape, dtype: np.empty(shape, dtype))
    monkeypatch.setattr(core, "_visible_gpus", lambda: [0])
    monkeypatch.setattr(core, "use_device", lambda d: __import__("contextlib").nullcontext())
    monkeypatch.setattr(core, "_to_device", lambda buf, device: buf)       # the device-resident leg of the default compression
    monkeypatch.setattr(core, "_own_stream", lambda device, index=0: __import__("contextlib").nullcontext())
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)


def _write_stack(folder, n, shape=(24, 40), dtype=np.uint16, seed=0):
    rng = np.random.default_rng(seed)
    folder.mkdir(parents=True, exist_ok=True)
    planes = {}
    for z in range(n):
        img = rng.integers(1, 4000, shape).astype(dtype)
        _io.write_tiff(folder / f"img_{z:04d}.tif", img, None)
        planes[f"img_{z:04d}.tif"] = img
    return planes


def test_batches_are_decoded_processed_and_written(tmp_path, host_only, monkeypatch):
    planes = _write_stack(tmp_path / "in", 11)
    seen = []

    def fake(stack, **kw):
        seen.append(stack.shape)
        assert kw["_max_batch"] == 4 and kw["sigma"] == (8, 8)
        return (stack // 2).astype(np.uint16)
    monkeypatch.setattr(core, "process_img", fake)
    rc = core.batch_filter(tmp_path / "in", tmp_path / "out", workers=4, threads_per_gpu=4, sigma=(8, 8), wavelet="db2")
    assert rc == 0
    assert sorted(s[0] for s in seen) == [3, 4, 4]                     # 11 planes in batches of 4
    for name, img in planes.items():
        assert np.array_equal(core.imread_tif_raw_png(tmp_path / "out" / name), img // 2)


def test_gpu_deflated_results_are_laid_out_as_tiff_files(tmp_path, host_only, monkeypatch):
    """the default output scheme ('ADOBE_DEFLATE', 1): the compute stage hands the writer zlib streams per strip
    (core.gpu_deflate on the device; here the same layout produced with zlib) and the native writer lays them out as TIFF."""
    import zlib
    planes = _write_stack(tmp_path / "in", 7, shape=(37, 40))
    monkeypatch.setattr(core, "process_img", lambda stack, **kw: (stack + 1).astype(np.uint16))
    monkeypatch.setattr(core, "_deflatable", lambda res: True)
    made = []

    def host_deflate(res):
        n, rows, cols = res.shape
        rps = 8                                                           # 5 strips, the last one short
        streams = [[zlib.compress(res[i, r:r + rps].tobytes(), 1) for r in range(0, rows, rps)] for i in range(n)]
        sizes = np.array([[len(s) for s in row] for row in streams], np.uint32)
        offsets = (np.cumsum(sizes.ravel(), dtype=np.uint64) - sizes.ravel()).reshape(sizes.shape)
        made.append(n)
        return core._DeflatedPlanes(np.frombuffer(b"".join(b"".join(row) for row in streams), np.uint8), offsets, sizes, rps,
                                    (rows, cols), np.dtype(np.uint16))
    monkeypatch.setattr(core, "gpu_deflate", host_deflate)
    assert core.batch_filter(tmp_path / "in", tmp_path / "out", workers=4, threads_per_gpu=4, sigma=(8, 8), wavelet="db2") == 0
    assert sum(made) == 7
    from PIL import Image
    for name, img in planes.items():
        assert _io.probe(tmp_path / "out" / name)[2].compression == 8
        assert np.array_equal(_io.read(tmp_path / "out" / name), img + 1)
        with Image.open(tmp_path / "out" / name) as im:
            assert np.array_equal(np.array(im), img + 1)
    d = host_deflate(np.stack([planes["img_0000.tif"], planes["img_0001.tif"]]))
    assert np.array_equal(d.inflate(1), planes["img_0001.tif"])
    # another compression level keeps the host encoder
    made.clear()
    assert core.batch_filter(tmp_path / "in", tmp_path / "out6", workers=4, threads_per_gpu=4, sigma=(8, 8), wavelet="db2",
                             compression=("ADOBE_DEFLATE", 6)) == 0
    assert not made and np.array_equal(_io.read(tmp_path / "out6" / "img_0003.tif"), planes["img_0003.tif"] + 1)


def test_png_stacks_take_the_batch_path(tmp_path, host_only, monkeypatch):
    """greyscale PNG tiles (core.py:209-210) are decoded by the native codec into the batch buffer like TIFF tiles."""
    from PIL import Image
    rng = np.random.default_rng(3)
    (tmp_path / "in").mkdir()
    planes = {}
    for z in range(5):
        img = rng.integers(0, 60000, (33, 47)).astype(np.uint16)
        Image.fromarray(img).save(tmp_path / "in" / f"p_{z:03d}.png")
        planes[f"p_{z:03d}.tif"] = img
    batches = []
    monkeypatch.setattr(core, "process_img", lambda stack, **kw: (batches.append(stack.shape[0]), stack)[1])
    slow = []
    monkeypatch.setattr(core._BatchPipeline, "_slow_file", lambda self, job: slow.append(job))
    assert core.batch_filter(tmp_path / "in", tmp_path / "out", workers=4, threads_per_gpu=4, sigma=(8, 8), wavelet="db2",
                             compression=None) == 0
    assert sum(batches) == 5 and not slow
    for name, img in planes.items():
        assert np.array_equal(_io.read(tmp_path / "out" / name), img)


def test_a_failing_batch_is_reported_and_the_rest_still_runs(tmp_path, host_only, monkeypatch):
    """MemoryError / AssertionError used to kill the feeder thread and batch_filter still returned 0."""
    _write_stack(tmp_path / "in", 9)
    calls = []

    def fake(stack, **kw):
        calls.append(stack.shape[0])
        if len(calls) == 2:
            raise MemoryError("cudaMalloc failed")
        if len(calls) == 3:
            raise AssertionError("clip level")
        return stack
    monkeypatch.setattr(core, "process_img", fake)
    rc = core.batch_filter(tmp_path / "in", tmp_path / "out", workers=2, threads_per_gpu=3, sigma=(8, 8), wavelet="db2")
    assert rc == 2 and calls == [3, 3, 3]
    assert len(list((tmp_path / "out").glob("*.tif"))) == 3           # the first batch was written, the others reported


def test_odd_shapes_and_unreadable_files_take_the_per_file_path(tmp_path, host_only, monkeypatch):
    planes = _write_stack(tmp_path / "in", 5)
    odd = np.arange(30 * 50, dtype=np.uint16).reshape(30, 50)
    _io.write_tiff(tmp_path / "in" / "img_0002.tif", odd, None)        # another shape in the middle of a batch
    (tmp_path / "in" / "img_0003.tif").write_bytes(b"garbage")        # undecodable: warned about by the reader, skipped
    monkeypatch.setattr(core, "NUM_RETRIES", 2)
    shapes = []

    def fake(stack, **kw):
        shapes.append(stack.shape)
        return stack
    monkeypatch.setattr(core, "process_img", fake)
    rc = core.batch_filter(tmp_path / "in", tmp_path / "out", workers=2, threads_per_gpu=8, sigma=(8, 8), wavelet="db2")
    assert rc == 0            # an undecodable input is warned about and skipped, as the reference does (core.py:1529-1537)
    assert (30, 50) in shapes and any(len(s) == 3 for s in shapes)
    assert np.array_equal(core.imread_tif_raw_png(tmp_path / "out" / "img_0002.tif"), odd)
    assert np.array_equal(core.imread_tif_raw_png(tmp_path / "out" / "img_0004.tif"), planes["img_0004.tif"])
    assert not (tmp_path / "out" / "img_0003.tif").exists()


def test_timeout_writes_a_zero_dummy_tile(tmp_path, host_only, monkeypatch):
    """core.py:1736-1750: zeros of `new_size or tile_size`, uint8 when convert_to_8bit."""
    _write_stack(tmp_path / "in", 4)
    real_read = _io.read

    def slow_read(path, out=None, threads=None):
        if str(path).endswith("img_0001.tif"):
            time.sleep(1.0)
        return real_read(path, out=out, threads=threads)
    monkeypatch.setattr(_io, "read", slow_read)
    monkeypatch.setattr(core, "process_img", lambda stack, **kw: (stack >> 8).astype(np.uint8))
    rc = core.batch_filter(tmp_path / "in", tmp_path / "out", workers=4, threads_per_gpu=4, sigma=(8, 8), wavelet="db2",
                           timeout=0.3, tile_size=(24, 40), d_type="uint16", convert_to_8bit=True)
    assert rc == 0
    dummy = core.imread_tif_raw_png(tmp_path / "out" / "img_0001.tif")
    assert dummy.shape == (24, 40) and dummy.dtype == np.uint8 and not dummy.any()
    assert core.imread_tif_raw_png(tmp_path / "out" / "img_0000.tif").dtype == np.uint8


@pytest.mark.timeout(120)
def test_a_dying_stage_does_not_deadlock_the_pipeline(tmp_path, host_only, monkeypatch):
    """the encoder dies on its first group while the bounded queues are full: the other stages must unblock and the call
    must return a non-zero code instead of hanging."""
    _write_stack(tmp_path / "in", 40)

    class Boom(BaseException):
        pass

    def die(*a, **k):
        raise Boom("encoder died")
    monkeypatch.setattr(core, "process_img", lambda stack, **kw: stack)
    monkeypatch.setattr(_io, "write_tiff_batch", die)
    monkeypatch.setenv("B200STRIPE_FILE_GROUP", "2")          # 20 groups through queues of depth 2
    rc = core.batch_filter(tmp_path / "in", tmp_path / "out", workers=2, threads_per_gpu=2, sigma=(8, 8), wavelet="db2",
                           compression=None)
    assert rc != 0


def test_queue_runner_replaces_a_timed_out_item_with_a_dummy(tmp_path):
    from multiprocessing import Queue
    args_q, prog_q = Queue(), Queue()
    out = tmp_path / "z.tif"
    args_q.put(dict(input_file=tmp_path / "in.tif", output_file=out, tile_size=(8, 9), new_size=None, convert_to_8bit=False,
                    seconds=5.0))
    r = core.MultiProcessQueueRunner(prog_q, args_q, fun=_sleepy, timeout=0.2)
    r.run()                                                            # in-process: exercises the run loop itself
    img = core.imread_tif_raw_png(out)
    assert img.shape == (8, 9) and img.dtype == np.uint16 and not img.any()
    assert prog_q.get(timeout=1) is True and prog_q.get(timeout=1) is False


def _sleepy(input_file=None, output_file=None, seconds=0.0, **kw):
    time.sleep(seconds)


# ---------------------------------------------------------------------------------------------------- plan cache
class _FakePlan:
    made = 0

    def __init__(self, ctx, params, dec_lo=None, flat=None):
        type(self).made += 1
        self.lock = threading.RLock()
        self._users = 0
        self.closed = False
        self.params = params

        class I:
            n_passes = 0
        self.info = I()

    def close(self):
        with self.lock:
            self.closed = True


def test_plan_cache_is_lru_and_never_closes_a_plan_in_use(monkeypatch):
    monkeypatch.setattr(_native, "Plan", _FakePlan)
    monkeypatch.setattr(_native, "context", lambda d: None)
    monkeypatch.setattr(core, "_plans", type(core._plans)())
    monkeypatch.setattr(core, "_MAX_CACHED_PLANS", 3)
    kw = dict(process=0, sigma=(0, 0), level=0, wavelet="db2", threshold=None, padding_mode="wrap", bidirectional=False,
              log1p=True)
    a = core._get_plan(0, (8, 8), _native.U16, _acquire=True, **kw)     # in use for the whole test
    b = core._get_plan(0, (9, 9), _native.U16, **kw)
    c = core._get_plan(0, (10, 10), _native.U16, **kw)
    assert core._get_plan(0, (9, 9), _native.U16, **kw) is b            # touch b: c is now the oldest idle plan
    d = core._get_plan(0, (11, 11), _native.U16, **kw)
    assert c.closed and not a.closed and not b.closed and not d.closed  # LRU among idle plans, the in-use plan survives
    e = core._get_plan(0, (12, 12), _native.U16, **kw)
    assert b.closed and not a.closed and not e.closed
    # a plan another thread is running (holds its lock) is skipped as well
    holder_has_it, release = threading.Event(), threading.Event()

    def hold():
        with d.lock:
            holder_has_it.set()
            release.wait(5)
    t = threading.Thread(target=hold)
    t.start()
    holder_has_it.wait(5)
    f = core._get_plan(0, (13, 13), _native.U16, **kw)
    assert not d.closed and e.closed and not f.closed
    release.set()
    t.join()
    core._release_plan(a)
    core.clear_plan_cache()
    assert a.closed and d.closed and f.closed and len(core._plans) == 0


def test_out_of_memory_retry_closes_only_idle_plans(monkeypatch):
    made = []

    class P(_FakePlan):
        def __init__(self, *a, **k):
            if len(made) == 2:
                made.append("oom")
                raise MemoryError("cudaMalloc failed")
            super().__init__(*a, **k)
            made.append(self)
    monkeypatch.setattr(_native, "Plan", P)
    monkeypatch.setattr(_native, "context", lambda d: None)
    monkeypatch.setattr(core, "_plans", type(core._plans)())
    kw = dict(process=0, sigma=(0, 0), level=0, wavelet="db2", threshold=None, padding_mode="wrap", bidirectional=False,
              log1p=True)
    busy = core._get_plan(0, (8, 8), _native.U16, _acquire=True, **kw)
    idle = core._get_plan(0, (9, 9), _native.U16, **kw)
    fresh = core._get_plan(0, (10, 10), _native.U16, **kw)              # first attempt raises, idle plans are released
    assert idle.closed and not busy.closed and not fresh.closed

"""GPU: the host layer's stream-ordered device frames under real threads.

Calls on CUDA-domain frames return before the GPU has finished; these tests
check the orderings the library promises: a frame written on one thread and read on another,
memory handed back to a CUDA domain while work on it is still queued, and many workers
hammering one domain -- always against the oracle, bit-exact."""
import ctypes
import threading

import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu
ORACLE = helpers.load_oracle()

W, H, DEPTH, FILT = 704, 576, 4, 0          # big enough that a transform outlives the call that enqueued it


def _coef_frames(compat, n, seed):
    rng = np.random.default_rng(seed)
    params = compat.make_params(W, H, FILT, DEPTH)
    iw, ih = params.iwt_luma_width, params.iwt_luma_height
    pinned = compat.pinned_domain()
    frames, wants = [], []
    for _ in range(n):
        f = compat.frame_new_and_alloc(pinned, compat.FORMAT_S16_420, iw, ih)
        want = []
        for c in range(3):
            v = compat.frame_plane(f, c)
            v[...] = rng.integers(-255, 256, size=v.shape)
            want.append(helpers.cpu_wavelet(ORACLE, "oracle", "inv", v.copy(), FILT, DEPTH))
        frames.append(f)
        wants.append(want)
    return params, pinned, frames, wants


def _run_threads(fns):
    errs = []

    def wrap(fn):
        def go():
            try:
                fn()
            except BaseException as e:       # noqa: BLE001 - report in the main thread
                errs.append(e)
        return go
    ths = [threading.Thread(target=wrap(fn)) for fn in fns]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    if errs:
        raise errs[0]


def test_frame_written_on_one_thread_read_on_another(cuda):
    """Thread A uploads + inverse-transforms a CUDA-domain frame and returns at once; thread B
    downloads it.  B's copy must wait for A's kernels (last-writer event of the frame)."""
    from schroedinger_b200 import compat, lib
    n = 12
    params, pinned, hosts, wants = _coef_frames(compat, n, 1)
    cuda_dom = compat.cuda_domain()
    iw, ih = params.iwt_luma_width, params.iwt_luma_height
    dev = [compat.frame_new_and_alloc(cuda_dom, compat.FORMAT_S16_420, iw, ih) for _ in range(n)]
    outs = [compat.frame_new_and_alloc(pinned, compat.FORMAT_S16_420, iw, ih) for _ in range(n)]
    ready = [threading.Event() for _ in range(n)]

    scratch = compat.frame_new_and_alloc(cuda_dom, compat.FORMAT_S16_420, iw, ih)

    def producer():
        for i in range(n):
            lib.schro_frame_to_gpu(dev[i], hosts[i])
            # a few milliseconds of queued work ahead of the transform B is going to wait for
            for _ in range(6):
                lib.schro_frame_iwt_transform(scratch, ctypes.byref(params))
                lib.schro_frame_inverse_iwt_transform(scratch, ctypes.byref(params))
            lib.schro_frame_inverse_iwt_transform(dev[i], ctypes.byref(params))   # in flight when we signal
            ready[i].set()
        lib.schro_b200_thread_release()

    def consumer():
        for i in range(n):
            ready[i].wait()
            lib.schro_gpuframe_to_cpu(outs[i], dev[i])
        lib.schro_b200_thread_release()

    _run_threads([producer, consumer])
    for i in range(n):
        for c in range(3):
            assert np.array_equal(compat.frame_plane(outs[i], c), wants[i][c]), (i, c)
    for f in dev + outs + hosts + [scratch]:
        lib.schro_frame_unref(f)
    lib.schro_memory_domain_free(cuda_dom)
    lib.schro_memory_domain_free(pinned)


def test_many_workers_share_one_cuda_domain(cuda):
    """Eight workers allocate, transform (the in-place call swaps regions and hands the old one
    back while the kernel that reads it is still queued), download and free frames of one
    CUDA domain; a block must never be reused while work on it is in flight."""
    from schroedinger_b200 import compat, lib
    nthreads, rounds = 8, 6
    params, pinned, hosts, wants = _coef_frames(compat, nthreads, 2)
    cuda_dom = compat.cuda_domain()
    iw, ih = params.iwt_luma_width, params.iwt_luma_height
    outs = [compat.frame_new_and_alloc(pinned, compat.FORMAT_S16_420, iw, ih) for _ in range(nthreads)]
    bad = []

    def worker(t):
        def go():
            for r in range(rounds):
                f = compat.frame_new_and_alloc(cuda_dom, compat.FORMAT_S16_420, iw, ih)
                lib.schro_frame_to_gpu(f, hosts[t])
                lib.schro_frame_inverse_iwt_transform(f, ctypes.byref(params))
                lib.schro_gpuframe_to_cpu(outs[t], f)
                lib.schro_frame_unref(f)
                for c in range(3):
                    if not np.array_equal(compat.frame_plane(outs[t], c), wants[t][c]):
                        bad.append((t, r, c))
            lib.schro_b200_thread_release()
        return go

    _run_threads([worker(t) for t in range(nthreads)])
    assert not bad, bad[:5]
    for f in outs + hosts:
        lib.schro_frame_unref(f)
    lib.schro_memory_domain_free(cuda_dom)
    lib.schro_memory_domain_free(pinned)


def test_block_matching_from_many_threads_matches_one_thread(cuda):
    """schro_hbm_scan + level-0 refinement for the same (picture, reference) pair from six
    threads at once (each on its own stream, the kernels on the high-priority side stream):
    every thread must get the single-threaded field."""
    from schroedinger_b200 import compat, lib
    w, h, levels = 352, 288, 3
    src, ref = helpers.panning_pair(w, h, np.random.default_rng(5))
    params = compat.make_params(w, h, xbsep=8, ybsep=8, xblen=12, yblen=12)
    oracle_fields, _, _ = helpers.oracle_hbm(ORACLE, src, ref, w, h, levels=levels)
    cuda_dom = compat.cuda_domain()

    def pyramid(img):
        frames = []
        cw, ch = w, h
        for l in range(levels + 1):
            f = compat.frame_new_and_alloc(cuda_dom, compat.FORMAT_U8_420, cw, ch, 32 if l == 0 else 8, 0)
            frames.append(f)
            cw, ch = (cw + 1) // 2, (ch + 1) // 2
        host = compat.frame_new_and_alloc(None, compat.FORMAT_U8_420, w, h, 32, 0)
        for c in range(3):
            compat.frame_plane(host, c)[...] = img[c]
        lib.schro_frame_to_gpu(frames[0], host)
        lib.schro_frame_unref(host)
        lib.schro_frame_mc_edgeextend(frames[0])
        for l in range(levels):
            lib.schro_frame_downsample(frames[l + 1], frames[l])
            lib.schro_frame_mc_edgeextend(frames[l + 1])
        return frames

    sp, rp = pyramid(src), pyramid(ref)
    arr = compat.FrameP * (levels + 1)
    n = params.x_num_blocks * params.y_num_blocks

    def field():
        hbm = lib.schro_hbm_new_from_frames(ctypes.byref(params), 0, levels, 0, arr(*sp), arr(*rp))
        lib.schro_hbm_scan(hbm)
        lib.schro_hierarchical_bm_scan_hint(hbm, 0, 3)
        mf = lib.schro_hbm_motion_field(hbm, 0)
        out = np.ctypeslib.as_array(ctypes.cast(mf.contents.motion_vectors, ctypes.POINTER(ctypes.c_uint8)),
                                    shape=(n * 20,)).copy()
        lib.schro_hbm_unref(hbm)
        return out

    want = field()
    for f in ("flags", "metric", "chroma_metric", "v"):
        assert np.array_equal(want.view(helpers.MV_DTYPE)[f], oracle_fields[0][f]), f
    got = [None] * 6

    def worker(t):
        def go():
            for _ in range(3):
                got[t] = field()
            lib.schro_b200_thread_release()
        return go

    _run_threads([worker(t) for t in range(6)])
    for t in range(6):
        assert np.array_equal(got[t], want), t
    for f in sp + rp:
        lib.schro_frame_unref(f)
    lib.schro_memory_domain_free(cuda_dom)


def test_block_matching_scan_on_one_thread_read_and_free_on_another(cuda):
    """SchroAsync moves the stages of a picture between workers: thread A enqueues the search (and
    returns without waiting -- the fields stay on the device), thread B reads level 0 and frees the
    object.  Host pyramid frames make every search allocate pooled device blocks on A that B hands
    back; far more than 64 rounds (the size of the per-thread pool this replaced)."""
    import queue
    from schroedinger_b200 import compat, lib
    w, h, levels = 176, 144, 2
    src, ref = helpers.panning_pair(w, h, np.random.default_rng(6), (2, -1))
    params = compat.make_params(w, h, xbsep=8, ybsep=8, xblen=12, yblen=12)
    want, _, _ = helpers.oracle_hbm(ORACLE, src, ref, w, h, levels=levels)

    def pyramid(img):
        frames = []
        cw, ch = w, h
        for l in range(levels + 1):
            frames.append(compat.frame_new_and_alloc(None, compat.FORMAT_U8_420, cw, ch, 32 if l == 0 else 8, 0))
            cw, ch = (cw + 1) // 2, (ch + 1) // 2
        for c in range(3):
            compat.frame_plane(frames[0], c)[...] = img[c]
        lib.schro_frame_mc_edgeextend(frames[0])
        for l in range(levels):
            lib.schro_frame_downsample(frames[l + 1], frames[l])
            lib.schro_frame_mc_edgeextend(frames[l + 1])
        return frames

    sp, rp = pyramid(src), pyramid(ref)
    arr = compat.FrameP * (levels + 1)
    n = params.x_num_blocks * params.y_num_blocks
    rounds = 150
    q = queue.Queue(maxsize=4)
    bad = []

    def producer():
        for _ in range(rounds):
            hbm = lib.schro_hbm_new_from_frames(ctypes.byref(params), 0, levels, 0, arr(*sp), arr(*rp))
            lib.schro_hbm_scan(hbm)
            lib.schro_hierarchical_bm_scan_hint(hbm, 0, 3)
            q.put(hbm)
        q.put(None)
        lib.schro_b200_thread_release()

    def consumer():
        k = 0
        while True:
            hbm = q.get()
            if hbm is None:
                break
            mf = lib.schro_hbm_motion_field(hbm, 0)
            got = np.ctypeslib.as_array(ctypes.cast(mf.contents.motion_vectors, ctypes.POINTER(ctypes.c_uint8)),
                                        shape=(n * 20,)).copy().view(helpers.MV_DTYPE)
            for f in ("flags", "metric", "v"):
                if not np.array_equal(got[f], want[0][f]):
                    bad.append((k, f))
            if k % 50 == 0:                      # now and then a coarser level too
                mf1 = lib.schro_hbm_motion_field(hbm, 1)
                g1 = np.ctypeslib.as_array(ctypes.cast(mf1.contents.motion_vectors, ctypes.POINTER(ctypes.c_uint8)),
                                           shape=(n * 20,)).copy().view(helpers.MV_DTYPE)
                if not np.array_equal(g1["v"], want[1]["v"]):
                    bad.append((k, "level 1"))
            lib.schro_hbm_unref(hbm)
            k += 1
        lib.schro_b200_thread_release()

    _run_threads([producer, consumer])
    assert not bad, bad[:5]
    for f in sp + rp:
        lib.schro_frame_unref(f)


def test_polling_wait_mode_and_thread_sync(cuda, monkeypatch):
    """SB2_HOST_WAIT_US switches the host waits from the driver's blocking wait to polling (read
    when a thread first enters the library); results are the same either way, and
    schro_b200_thread_sync drains what a thread has left in flight on CUDA-domain frames."""
    from schroedinger_b200 import compat, lib
    params, pinned, hosts, wants = _coef_frames(compat, 2, 5)
    cuda_dom = compat.cuda_domain()
    iw, ih = params.iwt_luma_width, params.iwt_luma_height
    outs = [compat.frame_new_and_alloc(pinned, compat.FORMAT_S16_420, iw, ih) for _ in range(2)]

    def worker(t):
        def go():
            f = compat.frame_new_and_alloc(cuda_dom, compat.FORMAT_S16_420, iw, ih)
            lib.schro_frame_to_gpu(f, hosts[t])
            lib.schro_frame_inverse_iwt_transform(f, ctypes.byref(params))   # left in flight
            lib.schro_b200_thread_sync()
            lib.schro_gpuframe_to_cpu(outs[t], f)
            lib.schro_frame_unref(f)
            lib.schro_b200_thread_release()
        return go

    for t, mode in enumerate(("20", "0")):
        monkeypatch.setenv("SB2_HOST_WAIT_US", mode)
        _run_threads([worker(t)])
        for c in range(3):
            assert np.array_equal(compat.frame_plane(outs[t], c), wants[t][c]), (mode, c)
    for f in outs + hosts:
        lib.schro_frame_unref(f)
    lib.schro_memory_domain_free(cuda_dom)
    lib.schro_memory_domain_free(pinned)


def test_in_place_transforms_with_the_spare_region(cuda):
    """The per-thread spare region of the in-place transforms (host core: sb2h_spare_take / _put): a worker that
    transforms its own persistent CUDA-domain frame over and over swaps between two regions without going
    through the domain; a frame that wanders between workers (uploaded on A, transformed on B, transformed back
    on C, downloaded on A) must not be taken as anyone's spare while another worker still has work on it."""
    from schroedinger_b200 import compat, lib
    nthreads, rounds = 6, 10
    params, pinned, hosts, wants = _coef_frames(compat, nthreads, 5)
    cuda_dom = compat.cuda_domain()
    iw, ih = params.iwt_luma_width, params.iwt_luma_height
    outs = [compat.frame_new_and_alloc(pinned, compat.FORMAT_S16_420, iw, ih) for _ in range(nthreads)]
    own = [compat.frame_new_and_alloc(cuda_dom, compat.FORMAT_S16_420, iw, ih) for _ in range(nthreads)]
    bad = []

    def worker(t):
        def go():
            for r in range(rounds):
                lib.schro_frame_to_gpu(own[t], hosts[t])
                lib.schro_frame_inverse_iwt_transform(own[t], ctypes.byref(params))
                if r % 3 == 2:                                   # and back again: forward of the inverse = the input
                    lib.schro_frame_iwt_transform(own[t], ctypes.byref(params))
                    lib.schro_frame_inverse_iwt_transform(own[t], ctypes.byref(params))
                lib.schro_gpuframe_to_cpu(outs[t], own[t])
                for c in range(3):
                    if not np.array_equal(compat.frame_plane(outs[t], c), wants[t][c]):
                        bad.append((t, r, c))
            lib.schro_b200_thread_release()
        return go

    _run_threads([worker(t) for t in range(nthreads)])
    assert not bad, bad[:5]
    # a wandering frame: every step on another thread, no waits in between except the final download
    wander = compat.frame_new_and_alloc(cuda_dom, compat.FORMAT_S16_420, iw, ih)
    steps = [lambda: lib.schro_frame_to_gpu(wander, hosts[0]),
             lambda: lib.schro_frame_inverse_iwt_transform(wander, ctypes.byref(params)),
             lambda: lib.schro_frame_iwt_transform(wander, ctypes.byref(params)),
             lambda: lib.schro_frame_inverse_iwt_transform(wander, ctypes.byref(params)),
             lambda: lib.schro_gpuframe_to_cpu(outs[0], wander)]
    for rnd in range(4):
        for c in range(3):
            compat.frame_plane(outs[0], c)[...] = 0
        for k, fn in enumerate(steps):
            # a fresh thread per step; in the first round the threads leave their contexts registered (with work
            # possibly still in flight), later rounds release them -- both states of "another thread" for the spare
            def step(fn=fn, keep=(rnd == 0 and k in (1, 2, 3))):
                fn()
                if not keep:
                    lib.schro_b200_thread_release()
            _run_threads([step])
        for c in range(3):
            assert np.array_equal(compat.frame_plane(outs[0], c), wants[0][c]), (rnd, c)
    for f in outs + hosts + own + [wander]:
        lib.schro_frame_unref(f)
    lib.schro_memory_domain_free(cuda_dom)
    lib.schro_memory_domain_free(pinned)

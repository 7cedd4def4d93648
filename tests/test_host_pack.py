"""avdf_host_pack (csrc/host_pack.cu): the host-side gather of a batch's raw stream arrays into staging memory.
Pure host code - runs without a GPU. Byte-exact against numpy on ragged, unaligned, empty and tiny spans, with
every thread count, and under concurrent callers (the streaming runner may run more than one packer thread)."""
import ctypes
import threading

import numpy as np
import pytest

from audio_visual_deepfake_detection_b200 import native


def gather(srcs, dst, offsets, n_threads):
    n = len(srcs)
    S = (ctypes.c_void_p * n)(*[a.ctypes.data for a in srcs])
    D = (ctypes.c_void_p * n)(*[dst.ctypes.data + int(o) for o in offsets])
    N = (ctypes.c_size_t * n)(*[a.nbytes for a in srcs])
    native.check(native.lib().avdf_host_pack(S, D, N, n, n_threads), "avdf_host_pack")


def make_case(seed, n_spans, max_bytes, misalign):
    rng = np.random.default_rng(seed)
    sizes = [int(x) for x in rng.integers(0, max_bytes, n_spans)]
    sizes[0] = 0                                                     # an empty span
    srcs = [rng.integers(0, 256, s + 3, dtype=np.uint8)[3 * (i % 2):][:s] for i, s in enumerate(sizes)]   # odd source alignment
    offsets, o = [], misalign
    for s in sizes:
        offsets.append(o)
        o += s + int(rng.integers(0, 5))                             # small gaps: the guard bytes must stay untouched
    return srcs, offsets, o + 8


@pytest.mark.parametrize("n_threads", [1, 2, 3, 8, 200])
@pytest.mark.parametrize("max_bytes,n_spans", [(300, 40), (70000, 25), (3 << 20, 9)])
def test_gather_is_byte_exact(n_threads, max_bytes, n_spans):
    srcs, offsets, total = make_case(max_bytes + n_threads, n_spans, max_bytes, misalign=5)
    dst = np.full(total, 0xA5, dtype=np.uint8)
    want = dst.copy()
    for a, o in zip(srcs, offsets):
        want[o:o + a.nbytes] = a
    gather(srcs, dst, offsets, n_threads)
    assert np.array_equal(dst, want)


def test_stream_shaped_batch_and_argument_errors():
    rng = np.random.default_rng(3)
    srcs = []
    for b in range(32):                                             # BYOL-A [T_b, 2048] and emotion2vec [T_e, 768] rows of one batch
        d = rng.uniform(4, 12)
        srcs.append(rng.standard_normal((int(12.497 * d), 2048)).astype(np.float32))
        srcs.append(rng.standard_normal((int(50 * d), 768)).astype(np.float32))
    offsets = np.cumsum([0] + [a.nbytes for a in srcs])
    dst = np.zeros(int(offsets[-1]), dtype=np.uint8)
    gather(srcs, dst, offsets[:-1], 4)
    for a, o in zip(srcs, offsets):
        assert np.array_equal(dst[o:o + a.nbytes].view(np.float32).reshape(a.shape), a)
    L = native.lib()
    assert L.avdf_host_pack(None, None, None, 0, 1) == 0            # nothing to do
    assert L.avdf_host_pack(None, None, None, 3, 1) == -1 and b"invalid argument" in L.avdf_last_error()
    one = (ctypes.c_void_p * 1)(None)
    assert L.avdf_host_pack(one, one, (ctypes.c_size_t * 1)(16), 1, 1) == -1
    assert L.avdf_host_pack(one, one, (ctypes.c_size_t * 1)(16), 1, 0) == -1


def test_concurrent_callers_do_not_mix_jobs():
    cases = [make_case(100 + i, 12, 1 << 20, misalign=i) for i in range(4)]
    dsts = [np.zeros(c[2], dtype=np.uint8) for c in cases]
    errs = []

    def work(i):
        try:
            for _ in range(6):
                dsts[i][:] = 0
                gather(cases[i][0], dsts[i], cases[i][1], 3)
                for a, o in zip(cases[i][0], cases[i][1]):
                    assert np.array_equal(dsts[i][o:o + a.nbytes], a)
        except BaseException as e:          # noqa: BLE001
            errs.append(e)
    ths = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert not errs, errs


def test_pageable_arrays_are_not_reported_pinned():
    """Ordinary numpy memory is never 'pinned' (also on a box without a GPU driver: the query fails -> 0 -> staging path)."""
    a = np.zeros(1 << 16, dtype=np.float32)
    S = (ctypes.c_void_p * 1)(a.ctypes.data)
    N = (ctypes.c_size_t * 1)(a.nbytes)
    L = native.lib()
    assert L.avdf_host_all_pinned(S, N, 1) == 0
    assert L.avdf_host_all_pinned(None, None, 0) == 0


def test_h2d_gather_rejects_bad_arguments_without_touching_cuda():
    L = native.lib()
    assert L.avdf_h2d_gather(None, None, None, 0, None) == 0                 # nothing to copy
    assert L.avdf_h2d_gather(None, None, None, 2, None) == -1 and b"invalid argument" in L.avdf_last_error()
    assert L.avdf_h2d_gather(None, None, None, -1, None) == -1
    one = (ctypes.c_void_p * 1)(None)
    assert L.avdf_h2d_gather(one, one, (ctypes.c_size_t * 1)(64), 1, None) == -1   # null span with a non-zero size
    assert L.avdf_h2d_gather(one, one, (ctypes.c_size_t * 1)(0), 1, None) == 0     # empty spans are skipped


def test_bf16_shard_is_round_to_nearest_even():
    """streaming.bf16_shard (the opt-in 16-bit feature-shard format) produces torch's bf16 bit patterns."""
    import torch
    from audio_visual_deepfake_detection_b200.libs.modeling.streaming import bf16_shard, collate_pinned
    rng = np.random.RandomState(0)
    x = (rng.standard_normal((257, 64)) * np.exp(rng.uniform(-20, 20, (257, 64)))).astype(np.float32)
    x[0, :4] = [0.0, -0.0, 1.00390625, 1.01171875]                 # exact ties: round to even
    want = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(bf16_shard(x), want)
    chunk = [{"video_id": "a", "duration": 5.0, "streams": {"byola": want[:10], "emo": want[10:30]}},
             {"video_id": "b", "duration": 6.0, "streams": {"byola": want[30:50], "emo": want[50:60]}}]
    out = collate_pinned(chunk)
    assert out[0]["streams"]["byola"].dtype == np.uint16 and np.array_equal(out[1]["streams"]["emo"], want[50:60])
    # stream-major, one block: every video's rows of a stream are adjacent in memory
    a0, a1 = out[0]["streams"]["byola"], out[1]["streams"]["byola"]
    assert a1.ctypes.data == a0.ctypes.data + a0.nbytes

"""Host logic of the index-free decompressor (pcodec_b200/csrc/host_common.hpp follow_chunk_chain, through the test hook
pco_b200_debug_follow_chain): which of the candidate chunk starts of a speculative walk are real chunks.  No device needed."""
import ctypes as C

import numpy as np

from pcodec_b200 import _lib

OK, CORRUPT = 0, 1


def _follow(cand, st, ends, pos, n0, out_off=0, dst_len=1 << 40, src_len=1 << 40):
    L = _lib.lib()
    L.pco_b200_debug_follow_chain.restype = C.c_size_t
    cand = np.ascontiguousarray(cand, dtype=np.uint64)
    st = np.ascontiguousarray(st, dtype=np.uint32)
    ends = np.ascontiguousarray(ends, dtype=np.uint64)
    ver = np.zeros(max(len(cand), 1), dtype=np.uint32)
    nxt = C.c_uint64()
    k = L.pco_b200_debug_follow_chain(cand.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p), ends.ctypes.data_as(C.c_void_p), C.c_uint32(len(cand)), C.c_uint64(pos),
                                      C.c_uint64(n0), C.c_uint64(out_off), C.c_uint64(dst_len), C.c_uint64(src_len), ver.ctypes.data_as(C.c_void_p), C.byref(nxt))
    return list(ver[:k]), nxt.value


def test_real_chunks_are_chained_and_coincidences_skipped():
    # real chunks at 11, 500, 990 (ends 500, 990, 1480); coincidences at 200 and 700 whose walks even "succeed"
    cand = [11, 200, 500, 700, 990]
    ends = [500, 333, 990, 999, 1480]
    ver, nxt = _follow(cand, [OK] * 5, ends, 11, 100)
    assert ver == [0, 2, 4] and nxt == 1480


def test_chain_stops_at_a_failed_walk_and_at_a_missing_successor():
    cand = [11, 500, 990, 1480]
    ver, nxt = _follow(cand, [OK, OK, CORRUPT, OK], [500, 990, 1480, 2000], 11, 100)
    assert ver == [0, 1] and nxt == 990  # the serial walker takes over at the failing chunk
    ver, nxt = _follow([11, 500, 990], [OK, OK, OK], [500, 777, 1480], 11, 100)
    assert ver == [0, 1] and nxt == 777  # chunk 1 ends where no candidate is (a chunk of another size follows)


def test_first_position_must_be_a_candidate():
    ver, nxt = _follow([20, 500], [OK, OK], [500, 900], 11, 100)
    assert ver == [] and nxt == 11


def test_destination_room_limits_the_chain():
    cand, ends = [11, 500, 990, 1480], [500, 990, 1480, 2000]
    ver, nxt = _follow(cand, [OK] * 4, ends, 11, 100, out_off=50, dst_len=260)
    assert ver == [0, 1] and nxt == 990  # 50 + 2 * 100 <= 260 < 50 + 3 * 100
    ver, nxt = _follow(cand, [OK] * 4, ends, 11, 100, out_off=0, dst_len=99)
    assert ver == [] and nxt == 11


def test_ends_must_advance_and_stay_inside_the_file():
    ver, nxt = _follow([11, 500], [OK, OK], [11, 900], 11, 100)  # a walk that claims to end where it began
    assert ver == [] and nxt == 11
    ver, nxt = _follow([11, 500], [OK, OK], [500, 5000], 11, 100, src_len=1000)  # an end behind the file
    assert ver == [0] and nxt == 500


def test_many_chunks_random_layout():
    rng = np.random.default_rng(0)
    sizes = rng.integers(50, 400, size=300)
    starts = np.concatenate([[17], 17 + np.cumsum(sizes)])[:-1]
    real_ends = starts + sizes
    fake = np.setdiff1d(rng.integers(17, int(real_ends[-1]), size=200), starts)
    cand = np.sort(np.concatenate([starts, fake]))
    is_real = np.isin(cand, starts)
    ends = np.where(is_real, 0, cand + rng.integers(1, 300, size=cand.size)).astype(np.uint64)
    ends[is_real] = real_ends
    # a fake's made-up end may by chance be another fake's start - it still is never reached from the real chain
    ver, nxt = _follow(cand, np.zeros(cand.size, dtype=np.uint32), ends, 17, 10)
    assert [int(cand[i]) for i in ver] == [int(x) for x in starts] and nxt == int(real_ends[-1])

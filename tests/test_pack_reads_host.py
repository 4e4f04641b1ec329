"""ngm_b200_pack_reads (host only): ASCII read rows -> 2 bit / base words + lengths + exception list, the staging format of
ngm_b200_run_batch (north_star: packed sequences in the staging buffers).  Checked against a plain numpy restatement of its contract
(include/ngm_b200.h): code A 0, C 1, G 2, T 3 (either case) at bits 2 i of word i / 16, length = last non-NUL byte + 1, every other byte
inside the length reads as code 0 and travels as {read, position, byte} in input order."""
import numpy as np
import pytest

CODE = {ord(c): v for v, cs in enumerate(("Aa", "Cc", "Gg", "Tt")) for c in cs}


def restate(reads):
    n, stride = reads.shape
    words = (stride + 15) // 16
    packed = np.zeros((n, words), np.uint32)
    lens = np.zeros(n, np.uint16)
    exc = []
    for r in range(n):
        nz = np.nonzero(reads[r])[0]
        ln = int(nz[-1]) + 1 if len(nz) else 0
        lens[r] = ln
        for i in range(ln):
            ch = int(reads[r, i])
            if ch in CODE:
                packed[r, i // 16] |= np.uint32(CODE[ch] << (2 * (i % 16)))
            else:
                exc.append((r, i, ch))
    return packed, lens, exc


@pytest.mark.parametrize("stride,threads", [(152, 1), (152, 3), (37, 2), (16, 1), (250, 4)])
def test_pack_reads_matches_the_contract(stride, threads):
    import ctypes as C
    from nextgenmap_b200.host.cuda_sw import READ_EXC, load_library
    rng = np.random.default_rng(stride * 7 + threads)
    n = 9000                                                   # (slices of at least 4096 reads per thread)
    reads = np.frombuffer(b"ACGTacgt", np.uint8)[rng.integers(0, 8, (n, stride))].copy()
    reads[rng.random((n, stride)) < 0.01] = ord("N")
    reads[rng.random((n, stride)) < 0.002] = ord("x")
    reads[rng.random((n, stride)) < 0.001] = 0                 # NUL inside a read: an exception like any other byte
    cut = rng.integers(0, stride + 1, n)
    for r in range(0, n, 3):
        reads[r, cut[r]:] = 0                                  # ragged lengths, empty rows
    reads[5] = 0
    reads[6, :] = ord("N")
    lib = load_library()                                       # host-only entry point: no device, no context
    row_bytes = 4 * ((stride + 15) // 16)
    packed, lens = np.zeros((n, row_bytes), np.uint8), np.zeros(n, np.uint16)
    exc, need = np.zeros(1 << 16, dtype=READ_EXC), C.c_size_t(0)
    rc = lib.ngm_b200_pack_reads(reads.ctypes.data, n, stride, packed.ctypes.data, row_bytes, lens.ctypes.data, exc.ctypes.data, len(exc), C.byref(need), threads)
    assert rc == n
    exc = exc[: need.value]
    want_p, want_l, want_e = restate(reads)
    np.testing.assert_array_equal(lens, want_l)
    np.testing.assert_array_equal(packed.view(np.uint32).reshape(n, -1)[:, : want_p.shape[1]], want_p)
    assert [(int(e["read_index"]), int(e["pos"]), int(e["ch"])) for e in exc] == want_e
    assert len(want_e) > 100

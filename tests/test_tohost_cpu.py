"""Host-side decoders of the nl_pairs_to_host transfer format (include/nlcuda.h): i from `first`, S from one-byte codes.
Pure host code of libnlcuda.so, so it is checked here without a GPU; the composed device -> host path is in test_tohost_gpu.py."""
import numpy as np
import pytest

import neighbourlists_jl_b200 as nl


@pytest.mark.parametrize("it,code", [(np.int32, 0), (np.int64, 1)])
def test_expand_rows_matches_repeat(it, code):
    L = nl._lib.lib()
    rng = np.random.default_rng(5)
    for trial in range(300):
        n = int(rng.integers(1, 80))
        cnt = rng.integers(0, 40 if trial % 4 == 0 else 9, n)
        if trial % 3 == 0:
            cnt[rng.integers(0, n, n // 2)] = 0          # empty rows, also at both ends
        first = np.concatenate([[1], 1 + np.cumsum(cnt)]).astype(it)
        P = int(first[-1] - 1)
        ref = np.repeat(np.arange(1, n + 1), cnt).astype(it)
        off = int(rng.integers(0, 4))                     # every alignment of the output within 16 bytes
        buf = np.full(P + 8, -7, dtype=it)
        out = buf[off:off + P]
        lo = int(rng.integers(0, P + 1))
        hi = int(rng.integers(lo, P + 1))
        if trial % 5 == 0:
            lo, hi = 0, P
        assert L.nl_host_expand_rows(code, first.ctypes.data, None, n, lo, hi, out.ctypes.data) == 0
        assert np.array_equal(out[lo:hi], ref[lo:hi])
        assert (buf[:off + lo] == -7).all() and (buf[off + hi:] == -7).all()   # nothing outside the range is touched
        # through a row -> global index map (shard lists)
        gmap = rng.integers(1, 10**6, n).astype(it)
        buf[:] = -7
        assert L.nl_host_expand_rows(code, first.ctypes.data, gmap.ctypes.data, n, lo, hi, out.ctypes.data) == 0
        assert np.array_equal(out[lo:hi], np.repeat(gmap, cnt)[lo:hi])
        assert (buf[:off + lo] == -7).all() and (buf[off + hi:] == -7).all()


@pytest.mark.parametrize("it,code", [(np.int32, 0), (np.int64, 1)])
def test_unpack_shifts_matches_definition(it, code):
    L = nl._lib.lib()
    rng = np.random.default_rng(6)
    for trial in range(300):
        P = int(rng.integers(1, 400))
        lo = int(rng.integers(0, P + 1))
        hi = int(rng.integers(lo, P + 1))
        if trial % 5 == 0:
            lo, hi = 0, P
        codes = rng.integers(0, 27, P).astype(np.uint8)
        if trial % 2:
            codes[rng.random(P) < 0.9] = 13              # mostly "no shift", like a real list
        c = codes.astype(np.int64)
        ref = np.stack([c % 3 - 1, (c // 3) % 3 - 1, c // 9 - 1], 1).astype(it)
        buf = np.full(3 * P + 16, -7, dtype=it)
        o = 4 * int(rng.integers(0, 2)) + (int(rng.integers(0, 4)) if trial % 7 == 0 else 0)   # aligned and unaligned bases
        S = buf[o:o + 3 * P].reshape(P, 3)
        assert L.nl_host_unpack_shifts(code, codes.ctypes.data, lo, hi, S.ctypes.data) == 0
        assert np.array_equal(S[lo:hi], ref[lo:hi])
        assert (buf[:o + 3 * lo] == -7).all() and (buf[o + 3 * hi:] == -7).all()


def test_decoder_argument_checks():
    L = nl._lib.lib()
    first = np.array([1, 3, 3, 6], dtype=np.int32)
    out = np.zeros(5, dtype=np.int32)
    assert L.nl_host_expand_rows(0, first.ctypes.data, None, 3, 0, 6, out.ctypes.data) == nl._lib.NL_ERR_BAD_ARG   # beyond first[n_rows] - 1
    assert L.nl_host_expand_rows(2, first.ctypes.data, None, 3, 0, 5, out.ctypes.data) == nl._lib.NL_ERR_BAD_ARG   # bad int_type
    assert L.nl_host_expand_rows(0, None, None, 3, 0, 5, out.ctypes.data) == nl._lib.NL_ERR_BAD_ARG
    assert L.nl_host_expand_rows(0, first.ctypes.data, None, 3, 2, 2, None) == 0                                     # empty range
    assert L.nl_host_unpack_shifts(0, None, 0, 4, out.ctypes.data) == nl._lib.NL_ERR_BAD_ARG
    assert L.nl_to_host_scratch_bytes(0) >= 256 and L.nl_to_host_scratch_bytes(1000) >= 1000 + 4
    # the whole-list entry point validates before touching CUDA
    p = nl._lib.NlParams()
    p.int_type = 0
    assert L.nl_pairs_to_host(p, None, 3, None, 5, None, None, None, None, 5, None, None, None, None, None, None, 0, 0, None) == nl._lib.NL_ERR_BAD_ARG


def test_decoders_sse2_path_in_a_subprocess():
    """The AVX2 variants are picked at run time; NL_HOST_NO_AVX2=1 forces the SSE2 ones, which must give the same arrays."""
    import os, subprocess, sys
    env = dict(os.environ, NL_HOST_NO_AVX2="1")
    r = subprocess.run([sys.executable, "-m", "pytest", __file__, "-q", "-x", "-k", "expand_rows or unpack_shifts"], env=env, capture_output=True, text=True,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]

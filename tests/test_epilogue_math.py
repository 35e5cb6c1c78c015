"""CPU checks of the arithmetic identities the tensor-core epilogues rely on (no GPU, no library
call): numpy float32 / float16 restatements of

  * the exact three-way FP16 split of the folded norm term (yb_knn.cu, write_row_extras),
  * the packed Hamming accumulator and its byte test (yb_knn_tf32.cu, process_group_ham;
    yb_hamming_tc.cu, k_ham_expand).

The GPU parity tests exercise the kernels themselves; these pin the maths they are built on."""
import numpy as np


def test_norm_term_splits_exactly_into_three_fp16_pieces():
    # beta = 2^(2 sigma - 1) |b|^2 / 2^15 (a float32) = h0 + h1 + h2 with every h an FP16 number;
    # the query side multiplies each piece by 2^15, the products are exact in FP32
    r = np.random.RandomState(0)
    for sigma in (-3, 0, 5, 12, 13):
        norms = np.concatenate([r.rand(2000) * 50, r.rand(2000) * 1e-3, [0.0, 1.0, 10.6875]]).astype(np.float32)
        sc = np.float32(2.0 ** sigma)
        w = norms * sc * sc * np.float32(0.5 / 32768.0)          # exact: powers of two
        keep = w <= 65504
        w = w[keep]
        h0 = w.astype(np.float16)
        r1 = w - h0.astype(np.float32)
        h1 = r1.astype(np.float16)
        r2 = r1 - h1.astype(np.float32)
        h2 = r2.astype(np.float16)
        total = h0.astype(np.float64) + h1.astype(np.float64) + h2.astype(np.float64)
        big = w >= 1.0   # below 1 the last bits fall under FP16's sub-normal grid (2^-24)
        assert np.array_equal(total[big], w[big].astype(np.float64))
        assert np.all(np.abs(total - w) <= 2.0 ** -24)
        # and the score identity: asc * (acc - 2^15 * (h0 + h1 + h2)) = |b|^2 - 2 <q, b>
        dot = r.rand(len(w)).astype(np.float64) * 3
        asc = -2.0 ** (1 - 2 * sigma)
        acc = dot * 2.0 ** (2 * sigma) - 32768.0 * total
        s = asc * acc
        want = norms[keep].astype(np.float64) - 2 * dot
        assert np.allclose(s[big], want[big], rtol=0, atol=1e-9 * np.maximum(1, np.abs(want[big])))


def _expand(codes_bits, scale):
    # bit set -> +scale, clear -> -scale (k_ham_expand)
    return np.where(codes_bits > 0, scale, -scale).astype(np.float64)


def test_packed_hamming_accumulator_and_byte_test():
    r = np.random.RandomState(1)
    bits = 64
    q = r.randint(0, 2, bits)
    rows = r.randint(0, 2, (3000, 3, bits))
    rows[0, 0] = q            # distance 0
    rows[1, 2] = 1 - q        # distance 64
    rows[2] = q               # all three slots at distance 0
    scales = (1.0, 16.0, 256.0)
    # one accumulator per combined row: slot i contributes scale_i^2 * <q, b_i>
    acc = np.zeros(len(rows))
    for i, sc in enumerate(scales):
        acc += (_expand(rows[:, i], sc) * _expand(q, sc)[None, :]).sum(1)
    ham = (rows != q[None, None, :]).sum(2)                      # [n][3]
    assert np.all(np.abs(acc) < 2 ** 23)                          # exact in FP32
    magic = np.float32(8388608.0 + 32 * 65793.0)
    y = (np.float32(-0.5) * acc.astype(np.float32) + magic).astype(np.float32)   # one FMA, exact
    u = y.view(np.uint32)
    for i in range(3):
        assert np.array_equal((u >> (8 * i)) & 0xff, ham[:, i].astype(np.uint32))
    assert np.all((u >> 24) == 0x4B)
    # "has a byte less than tau" on the three distance bytes: exact for the existence test
    for tau in (0, 1, 5, 17, 18, 33, 64, 65, 128):
        tau3 = np.uint32(tau * 0x010101)
        hit = (((u - tau3) & ~u) & np.uint32(0x808080)) != 0
        assert np.array_equal(hit, (ham < tau).any(1)), tau


def test_byte_test_exhaustive_for_two_byte_fields():
    # every pair of distances 0..128 against every threshold 0..128 (128-bit codes: two slots)
    a, b = np.meshgrid(np.arange(129, dtype=np.uint32), np.arange(129, dtype=np.uint32))
    x = (a | (b << 8) | np.uint32(0x4B000000)).ravel()
    for tau in range(129):
        tau3 = np.uint32(tau * 0x010101)
        hit = (((x - tau3) & ~x) & np.uint32(0x8080)) != 0
        want = ((x & 0xff) < tau) | (((x >> 8) & 0xff) < tau)
        assert np.array_equal(hit, want), tau

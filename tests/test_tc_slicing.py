"""CPU: the exact bf16 slicing of the tensor-core blocks (csrc/tc_block.cuh: slice3 / slice3_pair / magic_of), re-executed
in NumPy float32 arithmetic: three magic-number roundings split every f32 value x of a tile, on the grid 2^(E-8) with
2^E > max |x|, into p0 + p1 + p2 + r where every p_i is EXACTLY a bf16 number (the upper 16 bits of the f32 word),
|k_i| <= 256 grid units, |r| <= 2^(E-27) -- and the leading products of a K = 128 dot product stay below 2^24 grid units,
which is what makes their tensor-core accumulation exact whatever its rounding."""
import numpy as np
import pytest


def magic_of(bexp):
    bexp = min(max(int(bexp), 32), 230)
    m0b = np.uint32(((bexp + 16) << 23) | 0x400000)
    as_f = lambda u: np.array([u], dtype=np.uint32).view(np.float32)[0]
    return as_f(m0b), as_f(m0b - np.uint32(9 << 23)), as_f(m0b - np.uint32(18 << 23))


def slice3(x, m0, m1, m2):
    x = x.astype(np.float32)
    p0 = (x + m0) - m0
    r1 = x - p0
    p1 = (r1 + m1) - m1
    r2 = r1 - p1
    p2 = (r2 + m2) - m2
    return p0, p1, p2, r2 - p2


@pytest.mark.parametrize("scale_exp", [-20, -3, 0, 7])
def test_slices_are_exact_bf16_numbers_on_the_grid(scale_exp):
    rng = np.random.default_rng(scale_exp + 100)
    x = (rng.normal(size=4096) * 2.0 ** scale_exp).astype(np.float32)
    x[:4] = np.float32(0), np.float32(2.0 ** scale_exp), -np.abs(x).max(), np.abs(x).max()
    bexp = (np.abs(x).max().view(np.uint32) >> 23) & 0xFF      # exponent field of the tile maximum: 2^(bexp - 126) > max
    E = int(bexp) - 126
    m0, m1, m2 = magic_of(bexp)
    p0, p1, p2, r = slice3(x, m0, m1, m2)
    for i, p in enumerate((p0, p1, p2)):
        assert np.all(p.view(np.uint32) & 0xFFFF == 0), "a slice is not a bf16 number"
        k = p.astype(np.float64) / 2.0 ** (E - 8 - 9 * i)
        assert np.all(k == np.rint(k)) and np.abs(k).max() <= 256, (i, np.abs(k).max())
    # exact decomposition in f32 and the residual bound
    assert np.all((p0.astype(np.float64) + p1 + p2 + r) == x.astype(np.float64))
    assert np.abs(r).max() <= 2.0 ** (E - 27)
    # odd symmetry: the conjugate's slices are an exact sign flip (tc_rev.cuh slices conj(adjoint) this way)
    q0, q1, q2, _ = slice3(-x, m0, m1, m2)
    assert np.array_equal(q0, -p0) and np.array_equal(q1, -p1) and np.array_equal(q2, -p2)


def test_leading_products_of_a_block_row_accumulate_exactly():
    """K = 128 leading products of <= 2^16 grid units each: every partial sum is an integer below 2^24 on the common
    grid, i.e. exactly representable in the f32 accumulator, in any order and with any rounding mode."""
    rng = np.random.default_rng(5)
    w = rng.normal(size=128) / 8                               # a row of the real-ified 128 x 128 block
    w = np.clip(w, -1, 1)
    x = rng.normal(size=128).astype(np.float32)
    bexp = (np.abs(x).max().view(np.uint32) >> 23) & 0xFF
    E = int(bexp) - 126
    x0 = slice3(x, *magic_of(bexp))[0].astype(np.float64)
    w0 = np.rint(w / 2.0 ** -8) * 2.0 ** -8                     # host_slice3, first slice, grid 2^-8 for |w| <= 1
    unit = 2.0 ** (E - 8) * 2.0 ** -8
    prods = w0 * x0 / unit
    assert np.all(prods == np.rint(prods)) and np.abs(prods).max() <= 2 ** 16
    assert np.abs(prods).sum() < 2 ** 24                        # worst case 128 * 2^16 = 2^23
    acc = np.float32(0)
    for t in rng.permutation(128):
        acc = np.float32(acc + np.float32(w0[t] * x0[t]))
    assert float(acc) == float(np.sum(w0 * x0))

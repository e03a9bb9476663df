"""
oracle/philox.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Pure-Python / numpy Philox4x32-10 (Salmon et al., "Parallel random numbers: as
easy as 1, 2, 3", SC'11) and the "d3d stream v1" draw convention shared with
the device sampler (deconv3d_b200/csrc/d3d_rng.cuh):

    key     = (seed & 0xffffffff, seed >> 32)
    counter = (block, site, sweep, chain)        # four 32-bit words
    (x0,x1,x2,x3) = philox4x32_10(counter, key)
    draw 2*block   = ((x0 >> 5) * 2**26 + (x1 >> 6)) * 2**-53      in [0, 1)
    draw 2*block+1 = ((x2 >> 5) * 2**26 + (x3 >> 6)) * 2**-53

``site`` is the linear spaxel index y*W+x, ``sweep`` the reference's
``cur_iteration`` (0 = the initial-parameter draw, lib/run.py:310-314), and the
draws of one (chain, sweep, site) are consumed in the reference's call order
(lib/run.py:578 -> 435 -> lib/rtnorm.py:121-218): draws 0..P-1 are the jump
uniforms, draw P the acceptance uniform, draws P+1.. the truncated-normal
sub-stream.  The reference itself uses the unseeded global numpy MT19937
(lib/run.py:313,435,578; lib/rtnorm.py:17) and is not reproducible; this
stream replaces it on both sides of the parity comparison.
"""

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(counter, key):
    """counter: 4 ints, key: 2 ints (all < 2**32) -> 4 ints."""
    c0, c1, c2, c3 = [int(c) & MASK for c in counter]
    k0, k1 = [int(k) & MASK for k in key]
    for r in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> 32, p0 & MASK
        hi1, lo1 = p1 >> 32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def u53(hi32, lo32):
    """Two 32-bit words -> double in [0,1) with 53 random bits (same recipe as
    numpy's legacy ``random_sample``: (a>>5, b>>6))."""
    return ((hi32 >> 5) * 67108864.0 + (lo32 >> 6)) / 9007199254740992.0


def draw(seed, chain, sweep, site, k):
    """k-th uniform draw of (chain, sweep, site) of the d3d stream v1."""
    key = (seed & MASK, (seed >> 32) & MASK)
    x = philox4x32_10((k >> 1, site, sweep, chain), key)
    if k & 1:
        return u53(x[2], x[3])
    return u53(x[0], x[1])

// Philox4x32-10 (Salmon et al., SC'11) and the visiting order the Philox mode of the sweeps derives from it.  Plain
// functions, usable from host code as well (tests/test_philox_order.py compiles this header with g++ and checks the
// order on the CPU: bijection for every n, determinism, uniformity).
#pragma once
#if defined(__CUDACC__)
#define CELDA_HD __host__ __device__
#else
#define CELDA_HD
#endif

namespace celda {

// Counter-based generator.  Used for the production ("philox") mode of the sweeps; the injection mode takes order and
// uniforms from the host instead.
struct Philox4 {
    unsigned x[4];
};
CELDA_HD inline Philox4 philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
        const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1;
        const unsigned n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{{c0, c1, c2, c3}};
}

// sample(seq_len(n)) for one sweep (R/celda_C.R:640, R/celda_G.R:615) without a sort: the t-th visited item is
// philox_order(t), a keyed pseudo-random permutation of [0, n).  An alternating (unbalanced) Feistel network over
// `bits` = max(6, ceil(log2 n)) bits whose round function is Philox4x32-10 on the counter (right half, tag | round,
// sweep, 2 * stream) -- eight rounds, each a bijection of the bits-bit strings whatever the round function returns --
// followed by cycle walking: values >= n are pushed through the network again, which keeps the map a bijection of
// [0, n) (the domain is below 2 n for n > 32, so a value walks less than twice on average).  Every item is computed
// independently: no keys, no sort, no library kernels on the sweep path.
constexpr unsigned PHILOX_ORDER_ROUNDS = 8;
CELDA_HD inline unsigned philox_order_bits(unsigned long long n) {
    unsigned b = 6;
    while ((1ull << b) < n) ++b;
    return b;
}
CELDA_HD inline unsigned long long philox_order(unsigned long long t, unsigned long long n, unsigned bits, unsigned k0,
                                                unsigned k1, unsigned sweep, unsigned stream) {
    unsigned long long x = t;
    do {
        unsigned wl = bits >> 1, wr = bits - wl;  // widths of the left and right part; they swap every round
        unsigned long long L = x >> wr, R = x & ((1ull << wr) - 1ull);
        for (unsigned r = 0; r < PHILOX_ORDER_ROUNDS; ++r) {
            const Philox4 f = philox4x32_10((unsigned)R, 0x4f524400u | r, sweep, 2u * stream, k0, k1);
            const unsigned long long nr = L ^ ((unsigned long long)f.x[0] & ((1ull << wl) - 1ull));
            L = R;
            R = nr;
            const unsigned w = wl;
            wl = wr;
            wr = w;
        }
        x = (L << wr) | R;
    } while (x >= n);
    return x;
}

// the uniform of the t-th visited item: (52 random bits + 1/2) * 2^-52 in (0, 1), exactly representable
CELDA_HD inline double philox_uniform(unsigned long long t, unsigned k0, unsigned k1, unsigned sweep, unsigned stream) {
    const Philox4 b = philox4x32_10((unsigned)t, (unsigned)(t >> 32), sweep, 2u * stream + 1u, k0, k1);
    const unsigned long long bits = (((unsigned long long)b.x[0] << 32) | b.x[1]) >> 12;
    return ((double)bits + 0.5) * 2.220446049250313e-16;
}

}  // namespace celda

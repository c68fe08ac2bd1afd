// ptg_rng.cuh -- numpy's Generator(PCG64(SeedSequence(seed))).normal(), restated for the device.
//
// The reference draws its transition noise from gymnasium's Env.np_random (env/ptg_gym_env.py:584-585,598-599,
// 620-621), i.e. numpy's PCG64 bit generator + 256-layer ziggurat.  To keep state trajectories bit-exact without
// a host-drawn tape, the three pieces are restated here:
//   * SeedSequence(seed).generate_state(4, uint64)   (numpy/random/bit_generator.pyx, hashmix/mix pool of 4)
//   * PCG64: 128-bit LCG (setseq) + XSL-RR 64-bit output, pcg64_set_seed two-step initialisation
//   * random_standard_normal: ziggurat with the tables of ptg_ziggurat_tables.h (recovered from numpy and
//     verified bit-identical by tools/extract_numpy_ziggurat.py)
// tests/test_rng.py checks the host twins of these functions against numpy; the -m gpu tests check the device
// stream against numpy tapes.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PTG_HD __host__ __device__ __forceinline__
#else
#define PTG_HD inline
#endif

struct Pcg64 {
    uint64_t s_hi, s_lo;   // 128-bit LCG state
    uint64_t i_hi, i_lo;   // 128-bit increment (odd)
};

PTG_HD uint64_t ptg_mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

// state = state * 0x2360ED051FC65DA44385DF649FCCF645 + inc  (mod 2^128)
PTG_HD void pcg64_step(Pcg64& g) {
    const uint64_t m_hi = 0x2360ED051FC65DA4ull, m_lo = 0x4385DF649FCCF645ull;
    uint64_t lo = g.s_lo * m_lo;
    uint64_t hi = ptg_mulhi64(g.s_lo, m_lo) + g.s_hi * m_lo + g.s_lo * m_hi;
    uint64_t lo2 = lo + g.i_lo;
    hi += g.i_hi + (lo2 < lo ? 1ull : 0ull);
    g.s_lo = lo2;
    g.s_hi = hi;
}

// pcg64_next64: step, then XSL-RR of the new state
PTG_HD uint64_t pcg64_next64(Pcg64& g) {
    pcg64_step(g);
    uint64_t x = g.s_hi ^ g.s_lo;
    unsigned rot = (unsigned)(g.s_hi >> 58);
    return (x >> rot) | (x << ((64u - rot) & 63u));
}

PTG_HD double pcg64_next_double(Pcg64& g) { return (double)(pcg64_next64(g) >> 11) * (1.0 / 9007199254740992.0); }

// Generator(PCG64(SeedSequence(seed))) for a non-negative integer seed < 2^64
PTG_HD Pcg64 pcg64_from_seed(uint64_t seed) {
    const uint32_t INIT_A = 0x43b0d7e5u, MULT_A = 0x931e8875u, INIT_B = 0x8b51f9ddu, MULT_B = 0x58f38dedu,
                   MIX_L = 0xca01f9ddu, MIX_R = 0x4973f715u;
    uint32_t ent[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    int n_ent = (seed >> 32) ? 2 : 1;
    uint32_t hc = INIT_A, pool[4];
#define PTG_HASHMIX(v_in, out)                 \
    {                                          \
        uint32_t v_ = (v_in) ^ hc;             \
        hc *= MULT_A;                          \
        v_ *= hc;                              \
        v_ ^= v_ >> 16;                        \
        (out) = v_;                            \
    }
    for (int q = 0; q < 4; ++q) PTG_HASHMIX(q < n_ent ? ent[q] : 0u, pool[q]);
    for (int s = 0; s < 4; ++s)
        for (int d = 0; d < 4; ++d)
            if (s != d) {
                uint32_t h;
                PTG_HASHMIX(pool[s], h);
                uint32_t r = MIX_L * pool[d] - MIX_R * h;
                r ^= r >> 16;
                pool[d] = r;
            }
#undef PTG_HASHMIX
    uint32_t hb = INIT_B, st[8];
    for (int q = 0; q < 8; ++q) {
        uint32_t v = pool[q & 3] ^ hb;
        hb *= MULT_B;
        v *= hb;
        v ^= v >> 16;
        st[q] = v;
    }
    uint64_t w0 = st[0] | ((uint64_t)st[1] << 32), w1 = st[2] | ((uint64_t)st[3] << 32);
    uint64_t w2 = st[4] | ((uint64_t)st[5] << 32), w3 = st[6] | ((uint64_t)st[7] << 32);
    // pcg_setseq_128_srandom_r(initstate = (w0 << 64) | w1, initseq = (w2 << 64) | w3)
    Pcg64 g;
    g.i_hi = (w2 << 1) | (w3 >> 63);
    g.i_lo = (w3 << 1) | 1ull;
    g.s_hi = 0;
    g.s_lo = 0;
    pcg64_step(g);
    uint64_t lo = g.s_lo + w1;
    g.s_hi += w0 + (lo < g.s_lo ? 1ull : 0ull);
    g.s_lo = lo;
    pcg64_step(g);
    return g;
}

struct ZigTables {
    const uint64_t* ki;
    const double* wi;
    const double* fi;
    const uint64_t* kiwi;   // optional interleaved {ki[idx], bits(wi[idx])} pairs (e.g. a shared-memory copy), or null
};

// numpy random_standard_normal (ziggurat, 256 layers)
PTG_HD double pcg64_standard_normal(Pcg64& g, const ZigTables& z) {
    const double zig_r = 3.6541528853610087963519472518, zig_inv_r = 0.27366123732975827203338247596;
    for (;;) {
        uint64_t r = pcg64_next64(g);
        int idx = (int)(r & 0xff);
        r >>= 8;
        int sign = (int)(r & 0x1);
        uint64_t rabs = (r >> 1) & 0x000fffffffffffffull;
        uint64_t ki;
        double wi;
        if (z.kiwi) {
#if defined(__CUDA_ARCH__)
            const ulonglong2 kw = reinterpret_cast<const ulonglong2*>(z.kiwi)[idx];
            ki = kw.x; wi = __longlong_as_double((long long)kw.y);
#else
            ki = z.kiwi[2 * idx]; wi = z.wi[idx];
#endif
        } else {
            ki = z.ki[idx]; wi = z.wi[idx];
        }
        double x = (double)rabs * wi;
        if (sign) x = -x;
        if (rabs < ki) return x;   // 99.3 % of draws
        if (idx == 0) {
            for (;;) {
                double xx = -zig_inv_r * log1p(-pcg64_next_double(g));
                double yy = -log1p(-pcg64_next_double(g));
                if (yy + yy > xx * xx) return ((rabs >> 8) & 0x1) ? -(zig_r + xx) : zig_r + xx;
            }
        } else {
            if (((z.fi[idx - 1] - z.fi[idx]) * pcg64_next_double(g) + z.fi[idx]) < exp(-0.5 * x * x)) return x;
        }
    }
}

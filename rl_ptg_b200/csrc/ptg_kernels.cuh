// ptg_kernels.cuh -- CUDA kernels of the batched PtG environment (sm_100a).
//
//   k_build_*      one-off table construction on the device (window means, argmin LUT, market rows, clock)
//   k_construct    PTGEnv.__init__ for all envs
//   k_reset        VecEnv.reset() / PTGEnv.reset(seed)
//   k_step         VecEnv.step_wait(): T >= 1 env steps per launch, state in registers, SB3 auto-reset
//   k_episode_stats  deterministic reduction of finished-episode statistics (warp shuffles, last-CTA fold)
//
// The step kernel is HBM-bound streaming of per-env state + observation rows; the only gathers go to
// L2-resident tables (step table 64 B/entry, hour row 16*NV B, day row 32 B, argmin LUT 4 B) and -- on a noise
// draw -- to the env's 64 B RNG record.  The two [n_envs, price_ahead] observation blocks are transposed through
// shared memory per warp and leave the SM as one TMA bulk store each.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "ptg_device.cuh"

// ------------------------------------------------------------------------------------------------------------
// table construction
// ------------------------------------------------------------------------------------------------------------
struct BuildParams {
    const double* rows[PTG_N_DATASETS];   // device copies of the raw tables, row-major [len][7]
    int32_t len[PTG_N_DATASETS];
    int32_t ent_off[PTG_N_DATASETS];
    int32_t S;
    int32_t n_entries;
    const double* vals;                   // sorted distinct T values (incl. the reset temperature 16)
    int32_t n_vals;
    double t_cat_standby, t_cat_startup_cold, t_cat_startup_hot;
    double lo[6], hi[6];                  // normalisation bounds: T, h2, ch4, h2_res, h2o, heat
    RewardConsts rc;
};

// Virtual S-row window of _perform_sim_step (env/ptg_gym_env.py:525-557) starting at row s of table ds:
// rows [s, L) of the table, then either the last row repeated (change_operation == False) or the first rows of
// op1_start_p (the start-up tables, which are always stepped with change_operation == True, :386-388,:624).
struct WindowView {
    const double* head; int head_rows;      // rows s .. L-1 (may be 0)
    const double* tail; int tail_stride;    // tail rows: stride 7 (hand-over) or 0 (repeat last row)
    int n;                                  // total rows (== S except when the hand-over table is too short)
    __device__ double at(int q, int col) const {
        return q < head_rows ? head[(int64_t)q * 7 + col] : tail[(int64_t)(q - head_rows) * tail_stride + col];
    }
};

// numpy's pairwise summation (np.add.reduce) over column `col` of the window, rows [lo, lo + n)
__device__ double window_pairwise_sum(const WindowView& w, int col, int lo, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += w.at(lo + i, col);
        return res;
    } else if (n <= 128) {
        double r[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) r[q] = w.at(lo + q, col);
        int i;
        for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int q = 0; q < 8; ++q) r[q] += w.at(lo + i + q, col);
        }
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += w.at(lo + i, col);
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return window_pairwise_sum(w, col, lo, n2) + window_pairwise_sum(w, col, lo + n2, n - n2);
    }
}

__device__ __forceinline__ int find_val(const double* vals, int n, double v) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (vals[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;    // exact match guaranteed by construction
}

__device__ __forceinline__ int tinfo_of(const BuildParams& B, double T) {
    int flags = (T <= B.t_cat_startup_cold ? PTG_TF_COLD : 0) | (T >= B.t_cat_startup_hot ? PTG_TF_HOT : 0) |
                (T <= B.t_cat_standby ? PTG_TF_SBUP : 0);
    return (find_val(B.vals, B.n_vals, T) << 3) | flags;
}

__global__ void k_build_step_tab(const __grid_constant__ BuildParams B, StepEntry* out, StepMeans* out_means) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= B.n_entries) return;
    int ds = 0;
#pragma unroll
    for (int d = 1; d < PTG_N_DATASETS; ++d) if (g >= B.ent_off[d]) ds = d;
    const int s = g - B.ent_off[ds], L = B.len[ds], S = B.S;
    const bool handover = (ds == PTG_DS_STARTUP_COLD || ds == PTG_DS_STARTUP_HOT);
    WindowView w;
    w.head = B.rows[ds] + (int64_t)s * 7;
    w.head_rows = min(S, L - s);
    int ov = S - w.head_rows;                      // rows beyond the end of the table
    if (handover && w.head_rows > 0) {             // tail = next_operation[:overhead]
        w.tail = B.rows[PTG_DS_OP1_START_P];
        w.tail_stride = 7;
        ov = min(ov, B.len[PTG_DS_OP1_START_P]);
    } else {                                       // tail = operation[-1] repeated (also: s == L, any table)
        w.tail = B.rows[ds] + (int64_t)(L - 1) * 7;
        w.tail_stride = 0;
    }
    w.n = w.head_rows + ov;
    StepEntry e;
    StepMeans sm;
#pragma unroll
    for (int c = 0; c < 5; ++c) sm.mean[c] = window_pairwise_sum(w, 2 + c, 0, w.n) / (double)w.n;   // np.average
    sm.t_end = w.at(w.n - 1, 1);
    e.norm[0] = (float)((sm.t_end - B.lo[0]) / (B.hi[0] - B.lo[0]));
#pragma unroll
    for (int c = 0; c < 5; ++c) e.norm[1 + c] = (float)((sm.mean[c] - B.lo[1 + c]) / (B.hi[1 + c] - B.lo[1 + c]));
    reward_coefficients(B.rc, sm.mean, e.c_gas, e.c_eua, e.c_el, e.c_0);
    e.tinfo = tinfo_of(B, sm.t_end);
    e._pad = 0;
    out[g] = e;
    out_means[g] = sm;
}

// argmin LUT: lut[v][c] = first index of min |T_c[:] - vals[v]|  (np.abs(..).argmin(), :521-522)
__global__ void k_build_argmin(const __grid_constant__ BuildParams B, int32_t* lut) {
    const int v = blockIdx.x, c = blockIdx.y;
    const int targets[PTG_N_ARGMIN] = {PTG_DS_COOLDOWN, PTG_DS_STANDBY_UP, PTG_DS_STANDBY_DOWN,
                                       PTG_DS_STARTUP_COLD, PTG_DS_STARTUP_HOT, PTG_DS_OP1_START_P};
    const int ds = targets[c], L = B.len[ds];
    const double* tab = B.rows[ds];
    const double t = B.vals[v];
    double best = INFINITY;
    int besti = 0x7fffffff;
    for (int r = threadIdx.x; r < L; r += blockDim.x) {
        double d = fabs(tab[(int64_t)r * 7 + 1] - t);
        if (d < best) { best = d; besti = r; }      // ascending r per thread: first occurrence kept
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ob = __shfl_down_sync(0xffffffffu, best, o);
        int oi = __shfl_down_sync(0xffffffffu, besti, o);
        if (ob < best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    __shared__ double sb[32];
    __shared__ int si[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { sb[wid] = best; si[wid] = besti; }
    __syncthreads();
    if (wid == 0) {
        const int nw = blockDim.x >> 5;
        best = lane < nw ? sb[lane] : INFINITY;
        besti = lane < nw ? si[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ob = __shfl_down_sync(0xffffffffu, best, o);
            int oi = __shfl_down_sync(0xffffffffu, besti, o);
            if (ob < best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        if (lane == 0) lut[v * PTG_N_ARGMIN + c] = besti;
    }
}

// hour rows: [pa fp32 window | ... | packed part_full (2 bits each) | el fp64] in nv 16-byte vectors.
// Two layouts of the fp32 slots (the el price always sits in the last 8 bytes):
//   straight     window value a in slot a, the packed Part_Full codes in slot 4*nv - 3
//   interleaved  (PTG_HOUR_PAIRED, 64-byte rows of the key-major layout) first 32-byte half = even window values
//                a = 0, 2, ..., 12 + the packed codes, second half = odd values a = 1, 3, ..., 11 + el: a lane PAIR of the
//                step kernel gathers its two rows half by half (each LDG.256 touches 16 lines, not 32) and every lane
//                stages what it loaded -- value a of a row goes to column a of the staged tile, so even / odd lanes
//                fill even / odd columns without bank conflicts and without exchanging the window values
#ifndef PTG_HOUR_PAIRED
#define PTG_HOUR_PAIRED 1
#endif
#ifndef PTG_HOUR_QUAD
#define PTG_HOUR_QUAD 0          // 1: key-major mod layout, price_ahead 13: 128-byte hour rows with the day's prices, gathered by
#endif                           //    lane quads (no day-row gather).  Parity-green, but no faster: 52.7 vs 52.7 us uniform, 42.2 vs
                                 //    42.3 us sticky, and the roll-out kernel spills (32.9 vs 29.9 us) -- DESIGN.md section 7
#ifndef PTG_FLAT_PAIRED
#define PTG_FLAT_PAIRED 1        // flat layout (straight rows): pair gather as well (stage_flat_early_paired)
#endif
PTG_HD constexpr bool hour_interleaved(int nv, bool flat) { return PTG_HOUR_PAIRED && nv == 4 && !flat; }
PTG_HD constexpr int hour_slot(int nv, bool flat, int a) {           // fp32 slot of window value a
    return hour_interleaved(nv, flat) ? ((a & 1) ? 8 + (a >> 1) : (a >> 1)) : a;
}
PTG_HD constexpr int hour_bits_slot(int nv, bool flat) { return hour_interleaved(nv, flat) ? 7 : 4 * nv - 3; }

__global__ void k_build_hour_tab(const double* e_r_b, int n_hours, int pa, int nv, int raw, int flat, double lo, double hi,
                                 float* out, uint32_t* err) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_hours) return;
    float* row = out + (int64_t)t * nv * 4;
    const int src = raw ? 0 : 1;       // raw: el price window (:209) | mod: potential reward window (:208)
    uint32_t bits = 0;
    for (int a = 0; a < nv * 4 - 2; ++a) row[a] = 0.f;
    for (int a = 0; a < pa; ++a) {
        double v = e_r_b[((int64_t)src * pa + a) * n_hours + t];
        row[hour_slot(nv, flat != 0, a)] = (float)((v - lo) / (hi - lo));
        double pf = e_r_b[((int64_t)2 * pa + a) * n_hours + t];
        const bool ok = (pf == -1.0) || (pf == 0.0) || (pf == 1.0);
        if (!ok && !raw) atomicOr(err, PTG_EBIT_PARTFULL);
        const int code = ok ? (int)pf : 0;                 // 2-bit two's complement: -1 -> 0b11, 0 -> 0b00, 1 -> 0b01
        bits |= (uint32_t)(code & 3) << (2 * a);
    }
    row[hour_bits_slot(nv, flat != 0)] = __uint_as_float(bits);
    *reinterpret_cast<double*>(row + nv * 4 - 2) = e_r_b[t];          // e_r_b[0, 0, t]
}

__global__ void k_build_day_tab(const double* g_e, int n_days, double gas_lo, double gas_hi, double eua_lo,
                                double eua_hi, DayRow* out) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_days) return;
    DayRow r;
    r.gas = g_e[d];
    r.eua = g_e[(int64_t)2 * n_days + d];
    r.gas_n0 = (float)((g_e[d] - gas_lo) / (gas_hi - gas_lo));                              // :210
    r.gas_n1 = (float)((g_e[(int64_t)n_days + d] - gas_lo) / (gas_hi - gas_lo));
    r.eua_n0 = (float)((g_e[(int64_t)2 * n_days + d] - eua_lo) / (eua_hi - eua_lo));        // :211
    r.eua_n1 = (float)((g_e[(int64_t)3 * n_days + d] - eua_lo) / (eua_hi - eua_lo));
    out[d] = r;
}

// Quad hour rows (PTG_HOUR_QUAD; mod design, price_ahead == 13, key-major layout): ONE 128-byte line per hour that also
// carries the day's {gas, eua} -- for an env whose episode starts on a day boundary (act_ep_h == 24 * act_ep_d, true for
// every integer eps_len_d) the day index is t_hour / 24, so the day-row gather (32 more lines per warp-step on the
// L1TEX data pipe) disappears.  The row is four 32-byte quarters, one per lane of a lane QUAD (stage_windows_quad):
//   quarter q:  fp32 slots 0..3 = window values a = q, q + 4, q + 8, q + 12 (a < 13) | slot 4 = the packed Part_Full
//               codes of all 13 values (a copy per quarter) | slot 5 = 0 | fp64 in slots 6-7: q0 el, q1 gas, q2 eua, q3 0
__global__ void k_build_hourq_tab(const double* e_r_b, const double* g_e, int n_hours, int n_days, double lo, double hi,
                                  float* out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_hours) return;
    constexpr int pa = 13;
    float* row = out + (int64_t)t * 32;
    uint32_t bits = 0;
    for (int s = 0; s < 32; ++s) row[s] = 0.f;
    for (int a = 0; a < pa; ++a) {
        double v = e_r_b[((int64_t)1 * pa + a) * n_hours + t];                  // potential reward window (:208)
        row[8 * (a & 3) + (a >> 2)] = (float)((v - lo) / (hi - lo));
        double pf = e_r_b[((int64_t)2 * pa + a) * n_hours + t];
        const bool ok = (pf == -1.0) || (pf == 0.0) || (pf == 1.0);            // (k_build_hour_tab reports violations)
        bits |= (uint32_t)((ok ? (int)pf : 0) & 3) << (2 * a);
    }
    const int d = min(t / 24, n_days - 1);
    double* row_d = reinterpret_cast<double*>(row);
#pragma unroll
    for (int q = 0; q < 4; ++q) row[8 * q + 4] = __uint_as_float(bits);
    row_d[3] = e_r_b[t];                          // e_r_b[0, 0, t]
    row_d[7] = g_e[d];                            // g_e[0, 0, d]
    row_d[11] = g_e[(int64_t)2 * n_days + d];     // g_e[1, 0, d]
}

__global__ void k_build_clock_tab(int n, int sim_step, ClockRow* out) {
    int k1 = blockIdx.x * blockDim.x + threadIdx.x;
    if (k1 >= n) return;
    double clock_hours = (double)((long long)k1 * sim_step) / 3600;      // :442
    double clock_days = clock_hours / 24;
    ClockRow r;
    r.sin_h = (float)sin(2 * 3.141592653589793 * clock_hours);            // :449-450
    r.cos_h = (float)cos(2 * 3.141592653589793 * clock_hours);
    r.h_step = (int)floor(clock_hours);
    r.d_step = (int)floor(clock_days);
    out[k1] = r;
}

// ------------------------------------------------------------------------------------------------------------
// observation emission
// ------------------------------------------------------------------------------------------------------------
struct ObsRegs {            // what one env contributes to the observation, in registers
    int status;
    float norm[6];
    float sin_h, cos_h;
};

// --- TMA (bulk async copy) helpers: shared::cta -> global, 1-D ---------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_store_1d(void* gdst, const void* ssrc, uint32_t bytes) {
#if PTG_L2_HINTS
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes), "l"(PTG_L2_EVICT_FIRST) : "memory");
#else
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
#endif
}
// One lane of the (converged) warp, chosen by the hardware: unlike `lane == 0`, `elect.sync` tells ptxas that exactly
// one thread issues the bulk copies, so their operands move to uniform registers without a waterfall loop per copy
// (43 instead of 77 SASS instructions between the proxy fence and the bulk-group flush).
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk stores of this thread have finished READING shared memory (the staging buffer may be rewritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#define PTG_STAGE_FLOATS(NV) (32 * (4 * (NV) - 3))   // one [32 envs][pa <= 4*NV-3 values] staging region of a warp

// Observation emission.  The two [n_envs, price_ahead] blocks are transposed per warp through shared memory
// ([32][pa] floats, stride pa: conflict free for odd pa) and leave the SM as ONE bulk async copy (TMA) per block
// and warp -- 128*pa contiguous bytes -- issued by lane 0; the nine scalar keys are plain coalesced stores.
// PAC > 0 fixes price_ahead at compile time (the reference default 13) so the staging stores need no predicates.
template <int NV, bool MOD, int PAC>
__device__ __forceinline__ void stage_windows(const DevParams& P, float* sm, int lane, const float4 (&hrow)[NV],
                                              bool full_warp) {
    const int pa = PAC > 0 ? PAC : P.pa;
    // previous bulk stores (roll-out kernel, persistent CTAs) are done with sm.  Bulk groups belong to the thread that
    // issued them -- the elected lane; a ragged warp never issues any (and may get here from divergent code, where
    // elect.sync must not run)
    if (full_warp) { if (elect_one()) tma_store_wait_read(); }
    __syncwarp();
    float w[4 * NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) { w[4 * v] = hrow[v].x; w[4 * v + 1] = hrow[v].y; w[4 * v + 2] = hrow[v].z; w[4 * v + 3] = hrow[v].w; }
    float* rowA = sm + lane * pa;
#pragma unroll
    for (int a = 0; a < 4 * NV - 3; ++a)
        if (a < pa) rowA[a] = w[hour_slot(NV, false, a)];
    if (MOD) {
        // Part_Full (:239): 2-bit two's-complement codes (-1, 0, +1) -> sign-extending bit-field extract
        const int bits = __float_as_int(w[hour_bits_slot(NV, false)]);
        float* rowB = sm + PTG_STAGE_FLOATS(NV) + lane * pa;
#pragma unroll
        for (int a = 0; a < 16 && a < 4 * NV - 3; ++a)
            if (a < pa) rowB[a] = (float)((bits << (30 - 2 * a)) >> 30);
    }
}

// Full warps of the key-major layout with 64-byte interleaved hour rows (PTG_HOUR_PAIRED): lanes 2p and 2p + 1 gather
// their two rows TOGETHER -- the even lane the first 32-byte halves (even window values + Part_Full codes), the odd
// lane the second halves (odd window values + el price) -- so that each LDG.256 of the warp touches 16 lines instead
// of 32 (the L1TEX data pipe pays per instruction and 128-byte line: 64 -> 32 wavefronts per warp-step), and each lane
// stages what it holds: columns of its own parity in both rows of the pair (bank = 26p + 13r + 2c + parity: conflict
// free).  Exchanged between the lanes: the row index, the two Part_Full words and one el price.  Returns the lane's el.
template <bool MOD>
__device__ __forceinline__ double stage_windows_paired(const DevParams& P, float* sm, int lane, int t_hour) {
    constexpr int pa = 13;
    if (elect_one()) tma_store_wait_read();       // (as in stage_windows: previous bulk stores are done with sm)
    __syncwarp();
    const int t_p = __shfl_xor_sync(0xffffffffu, t_hour, 1);
    const int odd = lane & 1;
    const char* base = reinterpret_cast<const char*>(P.hour_tab) + 32 * odd;
    const U256 h0 = ldg256_nc(base + (int64_t)(odd ? t_p : t_hour) * 64);      // half `odd` of the even lane's row
    const U256 h1 = ldg256_nc(base + (int64_t)(odd ? t_hour : t_p) * 64);      // half `odd` of the odd lane's row
    float* rowA = sm + (lane - odd) * pa + odd;                                // row 2p, columns of this lane's parity
    const unsigned long long q0[4] = {h0.a, h0.b, h0.c, h0.d}, q1[4] = {h1.a, h1.b, h1.c, h1.d};
#pragma unroll
    for (int c = 0; c < 6; ++c) {                 // slots 0..5 of the half: a = 2c + odd
        rowA[2 * c] = __uint_as_float((uint32_t)(q0[c >> 1] >> (32 * (c & 1))));
        rowA[pa + 2 * c] = __uint_as_float((uint32_t)(q1[c >> 1] >> (32 * (c & 1))));
    }
    if (!odd) {                                   // slot 6 of the first half: a = 12 (the second half has el there)
        rowA[12] = __uint_as_float((uint32_t)h0.d);
        rowA[pa + 12] = __uint_as_float((uint32_t)h1.d);
    }
    if (MOD) {
        // Part_Full codes of both rows live in slot 7 of the first halves (even lane); the odd lane gets a copy, then
        // every lane decodes the columns of its parity: a = 2c + odd -> pre-shift by 2 * odd, compile-time shifts after
        const int b0 = __shfl_sync(0xffffffffu, (int)(h0.d >> 32), lane & ~1) >> (2 * odd);
        const int b1 = __shfl_sync(0xffffffffu, (int)(h1.d >> 32), lane & ~1) >> (2 * odd);
        float* rowB = rowA + PTG_STAGE_FLOATS(4);
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            rowB[2 * c] = (float)((b0 << (30 - 4 * c)) >> 30);
            rowB[pa + 2 * c] = (float)((b1 << (30 - 4 * c)) >> 30);
        }
        if (!odd) {
            rowB[12] = (float)((b0 << 6) >> 30);
            rowB[pa + 12] = (float)((b1 << 6) >> 30);
        }
    }
    // el price (last 8 bytes of the second half): the odd lane holds its own in h1 and its partner's in h0
    const unsigned long long el_p = __shfl_xor_sync(0xffffffffu, h0.d, 1);
    return __longlong_as_double((long long)(odd ? h1.d : el_p));
}

// Full warps whose lanes all sit on a day-aligned clock (t_day == t_hour / 24): the 128-byte quad rows of
// k_build_hourq_tab.  Lanes 4p .. 4p + 3 gather their four rows TOGETHER -- lane q the q-th 32-byte quarter of each --
// so every LDG.256 of the warp touches 8 lines and the four of them 32: what the 64-byte rows cost, with the day's gas
// and EUA price on board (no day-row gather: 64 -> 32 lines per warp-step for the market data).  Each lane stages the
// columns a = q (mod 4) of the four rows (bank = 20p + 13r + 4c + q: conflict free) and decodes the Part_Full codes of
// the same columns from its quarter's copy of the code word.  The three prices of a row arrive in three different
// lanes; they change hands through 768 bytes of the warp's (not yet staged) Part_Full tile.  Returns the lane's el.
__device__ __forceinline__ double stage_windows_quad(const DevParams& P, float* sm, int lane, int t_hour, double& gas,
                                                     double& eua) {
    constexpr int pa = 13;
    if (elect_one()) tma_store_wait_read();       // (as in stage_windows: previous bulk stores are done with sm)
    __syncwarp();
    const int q = lane & 3, l0 = lane & ~3;
    const char* base = reinterpret_cast<const char*>(P.hourq_tab) + 32 * q;
    U256 h[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int t_r = __shfl_sync(0xffffffffu, t_hour, l0 + r);
        h[r] = ldg256_nc(base + (int64_t)t_r * 128);                            // quarter q of the row of lane 4p + r
    }
    double* xb = reinterpret_cast<double*>(sm + PTG_STAGE_FLOATS(4));           // [32 rows][el, gas, eua]
    if (q < 3) {
#pragma unroll
        for (int r = 0; r < 4; ++r) xb[(l0 + r) * 3 + q] = __longlong_as_double((long long)h[r].d);
    }
    float* rowA = sm + l0 * pa + q;                                             // row 4p, column q
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        rowA[r * pa] = __uint_as_float((uint32_t)h[r].a);
        rowA[r * pa + 4] = __uint_as_float((uint32_t)(h[r].a >> 32));
        rowA[r * pa + 8] = __uint_as_float((uint32_t)h[r].b);
    }
    if (q == 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) rowA[r * pa + 12] = __uint_as_float((uint32_t)(h[r].b >> 32));
    }
    __syncwarp();
    const double el = xb[lane * 3];
    gas = xb[lane * 3 + 1];
    eua = xb[lane * 3 + 2];
    __syncwarp();                                 // the prices are out: the Part_Full tile may be staged over them
    float* rowB = rowA + PTG_STAGE_FLOATS(4);
#pragma unroll
    for (int r = 0; r < 4; ++r) {                 // codes of a = 4c + q: pre-shift by 2q, compile-time shifts after
        const int b = (int)(uint32_t)h[r].c >> (2 * q);
        rowB[r * pa] = (float)((b << 30) >> 30);
        rowB[r * pa + 4] = (float)((b << 22) >> 30);
        rowB[r * pa + 8] = (float)((b << 14) >> 30);
        if (q == 0) rowB[r * pa + 12] = (float)((b << 6) >> 30);
    }
    return el;
}

// The warp's two staged window tiles leave the SM as one bulk async copy (TMA) each.
template <int NV, bool MOD, int PAC>
__device__ __forceinline__ void issue_window_stores(const DevParams& P, float* __restrict__ obs, float* sm, int lane,
                                                    int warp_env0, int nvalid) {
    const int pa = PAC > 0 ? PAC : P.pa;
    float* smA = sm;
    float* smB = sm + PTG_STAGE_FLOATS(NV);
    float* gA = obs + P.off_win0 + warp_env0 * pa;
    float* gB = obs + P.off_win1 + warp_env0 * pa;
    if (nvalid == 32) {
        fence_proxy_async_smem();                 // make the generic-proxy writes visible to the async proxy
        __syncwarp();
        if (elect_one()) {
            tma_store_1d(gA, smA, 128u * (uint32_t)pa);
            if (MOD) tma_store_1d(gB, smB, 128u * (uint32_t)pa);
            tma_store_commit();
        }
    } else {                                      // ragged last warp: plain stores
        __syncwarp();
        for (int idx = lane; idx < nvalid * pa; idx += 32) {
            gA[idx] = smA[idx];
            if (MOD) gB[idx] = smB[idx];
        }
    }
}

// The nine scalar observation keys (+ the raw design's daily price pairs): plain coalesced stores.
template <bool MOD>
__device__ __forceinline__ void store_obs_scalars(const DevParams& P, float* __restrict__ obs, int e,
                                                  const DayRow& day, const ObsRegs& o) {
    if (!MOD) {
        st_stream(reinterpret_cast<float2*>(obs + P.off_gas) + e, make_float2(day.gas_n0, day.gas_n1));
        st_stream(reinterpret_cast<float2*>(obs + P.off_eua) + e, make_float2(day.eua_n0, day.eua_n1));
    }
    float* sc = obs + P.off_scalar + e;
    st_stream(sc, __int_as_float(o.status));
#pragma unroll
    for (int q = 0; q < 6; ++q) st_stream(sc + (1 + q) * P.n_pad, o.norm[q]);
    st_stream(sc + 7 * P.n_pad, o.sin_h);
    st_stream(sc + 8 * P.n_pad, o.cos_h);
}

template <int NV, bool MOD, int PAC>
__device__ __forceinline__ void flush_obs(const DevParams& P, float* __restrict__ obs, float* sm, int e,
                                          bool active, int lane, int warp_env0, int nvalid, const DayRow& day,
                                          const ObsRegs& o) {
    issue_window_stores<NV, MOD, PAC>(P, obs, sm, lane, warp_env0, nvalid);
    if (active) store_obs_scalars<MOD>(P, obs, e, day, o);
}

template <int NV, bool MOD>
__device__ __forceinline__ void emit_obs(const DevParams& P, float* __restrict__ obs, float* sm, int e,
                                         bool active, int lane, int warp_env0, int nvalid,
                                         const float4 (&hrow)[NV], const DayRow& day, const ObsRegs& o) {
    stage_windows<NV, MOD, 0>(P, sm, lane, hrow, nvalid == 32);
    flush_obs<NV, MOD, 0>(P, obs, sm, e, active, lane, warp_env0, nvalid, day, o);
}

// Everything one env contributes to an observation can be re-derived from five integers: the step-table entry
// (or -1 for the reset row), the market row indices, METH_STATUS and the clock index.  The cold paths below
// (terminal observation, masked reset, eval-mode info) take those scalars and re-read the L2-resident tables, so
// the hot path never has to keep addressable copies of its registers.
struct ObsKey {
    int ent, t_hour, t_day, status, k1;
};

// Scalar (uncoalesced) emission of one env's observation -- terminal observations and masked resets (rare).
template <int NV, bool MOD>
__device__ __noinline__ void emit_obs_scalar(const DevParams& P, float* __restrict__ obs, int64_t e, ObsKey key) {
    const float* w = reinterpret_cast<const float*>(P.hour_tab + (int64_t)key.t_hour * NV);
    for (int a = 0; a < P.pa; ++a) obs[P.off_win0 + e * P.pa + a] = w[hour_slot(NV, false, a)];
    if (MOD) {
        const uint32_t bits = __float_as_uint(w[hour_bits_slot(NV, false)]);
        for (int a = 0; a < P.pa; ++a) obs[P.off_win1 + e * P.pa + a] = (float)(((int)bits << (30 - 2 * a)) >> 30);
    } else {
        const DayRow day = P.day_tab[key.t_day];
        obs[P.off_gas + 2 * e] = day.gas_n0; obs[P.off_gas + 2 * e + 1] = day.gas_n1;
        obs[P.off_eua + 2 * e] = day.eua_n0; obs[P.off_eua + 2 * e + 1] = day.eua_n1;
    }
    float* sc = obs + P.off_scalar + e;
    sc[0] = __int_as_float(key.status);
    for (int q = 0; q < 6; ++q) sc[(1 + q) * P.n_pad] = key.ent >= 0 ? P.step_tab[key.ent].norm[q] : P.reset_norm[q];
    const ClockRow c = P.clock_tab[key.k1];
    sc[7 * P.n_pad] = c.sin_h;
    sc[8 * P.n_pad] = c.cos_h;
}

// ------------------------------------------------------------------------------------------------------------
// flat observation layout (PtgConfig.obs_layout == 1, price_ahead == 13): one [F]-float feature row per env in the
// column order of SB3's CombinedExtractor (ptg_features), written by the step kernel itself
//   mod (F = 40): CH4, heat, H2O, H2_in, H2_res | one-hot(status)[6] | Part_Full[13] | Pot_Reward[13] | T, cos, sin
//   raw (F = 31): CH4 | EUA[2] | heat | Elec_Price[13] | Gas[2] | H2O, H2_in, H2_res | one-hot[6] | T, cos, sin
// ------------------------------------------------------------------------------------------------------------
#define PTG_FLAT_F(MOD) ((MOD) ? 40 : 31)

// Cold path: one env's row with plain stores from re-read tables (reset kernel, terminal observations).
template <bool MOD>
__device__ __noinline__ void emit_flat_row(const DevParams& P, float* __restrict__ obs, int64_t e, ObsKey key) {
    float* r = obs + e * PTG_FLAT_F(MOD);
    const float* w = reinterpret_cast<const float*>(P.hour_tab + (int64_t)key.t_hour * 4);
    float nrm[6];
    for (int q = 0; q < 6; ++q) nrm[q] = key.ent >= 0 ? P.step_tab[key.ent].norm[q] : P.reset_norm[q];
    const ClockRow c = P.clock_tab[key.k1];
    int oh;
    if (MOD) {
        r[0] = nrm[2]; r[1] = nrm[5]; r[2] = nrm[4]; r[3] = nrm[1]; r[4] = nrm[3]; oh = 5;
        const int bits = __float_as_int(w[13]);
        for (int a = 0; a < 13; ++a) { r[11 + a] = (float)((bits << (30 - 2 * a)) >> 30); r[24 + a] = w[a]; }
        r[37] = nrm[0]; r[38] = c.cos_h; r[39] = c.sin_h;
    } else {
        const DayRow day = P.day_tab[key.t_day];
        r[0] = nrm[2]; r[1] = day.eua_n0; r[2] = day.eua_n1; r[3] = nrm[5];
        for (int a = 0; a < 13; ++a) r[4 + a] = w[a];
        r[17] = day.gas_n0; r[18] = day.gas_n1; r[19] = nrm[4]; r[20] = nrm[1]; r[21] = nrm[3]; oh = 22;
        r[28] = nrm[0]; r[29] = c.cos_h; r[30] = c.sin_h;
    }
    for (int q = 0; q < 6; ++q) r[oh + q] = key.status == q ? 1.0f : 0.0f;
}

// Hot path, early half: everything of the row that only needs the market rows.  mod: six 16-byte groups (Part_Full
// 1..12, Pot_Reward 0..11) as STS.128 (row stride 160 B: 2-way conflicts, 6 instructions instead of 24); the two
// window values that share a group with late scalars are returned.  raw: 17 conflict-free scalar stores (stride 31).
template <bool MOD>
__device__ __forceinline__ void stage_flat_early(float* row, const float4 (&h)[4], const DayRow& day, float& pf0, float& pr12) {
    if (MOD) {
        const int bits = __float_as_int(h[3].y);
        auto pf = [bits](int a) { return (float)((bits << (30 - 2 * a)) >> 30); };
        float4* r4 = reinterpret_cast<float4*>(row);
        r4[3] = make_float4(pf(1), pf(2), pf(3), pf(4));
        r4[4] = make_float4(pf(5), pf(6), pf(7), pf(8));
        r4[5] = make_float4(pf(9), pf(10), pf(11), pf(12));
        r4[6] = h[0]; r4[7] = h[1]; r4[8] = h[2];
        pf0 = pf(0); pr12 = h[3].x;
    } else {
        row[1] = day.eua_n0; row[2] = day.eua_n1;
        row[4] = h[0].x; row[5] = h[0].y; row[6] = h[0].z; row[7] = h[0].w; row[8] = h[1].x; row[9] = h[1].y;
        row[10] = h[1].z; row[11] = h[1].w; row[12] = h[2].x; row[13] = h[2].y; row[14] = h[2].z; row[15] = h[2].w;
        row[16] = h[3].x; row[17] = day.gas_n0; row[18] = day.gas_n1;
    }
}
// late half: plant scalars, one-hot status, clock
template <bool MOD>
__device__ __forceinline__ void stage_flat_late(float* row, const ObsRegs& o, float pf0, float pr12) {
    const int s = o.status;
    if (MOD) {
        float4* r4 = reinterpret_cast<float4*>(row);
        r4[0] = make_float4(o.norm[2], o.norm[5], o.norm[4], o.norm[1]);
        r4[1] = make_float4(o.norm[3], s == 0 ? 1.f : 0.f, s == 1 ? 1.f : 0.f, s == 2 ? 1.f : 0.f);
        r4[2] = make_float4(s == 3 ? 1.f : 0.f, s == 4 ? 1.f : 0.f, s == 5 ? 1.f : 0.f, pf0);
        r4[9] = make_float4(pr12, o.norm[0], o.cos_h, o.sin_h);
    } else {
        row[0] = o.norm[2]; row[3] = o.norm[5]; row[19] = o.norm[4]; row[20] = o.norm[1]; row[21] = o.norm[3];
#pragma unroll
        for (int q = 0; q < 6; ++q) row[22 + q] = s == q ? 1.f : 0.f;
        row[28] = o.norm[0]; row[29] = o.cos_h; row[30] = o.sin_h;
    }
}

__device__ __forceinline__ float4 f4_lo(unsigned long long a, unsigned long long b) {
    return make_float4(__uint_as_float((uint32_t)a), __uint_as_float((uint32_t)(a >> 32)), __uint_as_float((uint32_t)b),
                       __uint_as_float((uint32_t)(b >> 32)));
}

// Flat layout, mod design, full warps: the 64-byte (straight) hour rows of lanes 2p and 2p + 1 are gathered by the pair
// -- even lane: Pot_Reward 0..7 of both rows, odd lane: Pot_Reward 8..12, the Part_Full word and the el price of both
// (each LDG.256 touches 16 lines instead of 32, see stage_windows_paired) -- and each lane stages the Pot_Reward groups
// it holds for BOTH rows; the Part_Full groups stay with the row's own lane, which gets {Pot_Reward[12], codes} and el
// from the odd lane with two 64-bit shuffles.  Returns the lane's el price.
__device__ __forceinline__ double stage_flat_early_paired(const DevParams& P, float* sm, int lane, int t_hour, float& pf0,
                                                          float& pr12) {
    const int t_p = __shfl_xor_sync(0xffffffffu, t_hour, 1);
    const int odd = lane & 1;
    const char* base = reinterpret_cast<const char*>(P.hour_tab) + 32 * odd;
    const U256 h0 = ldg256_nc(base + (int64_t)(odd ? t_p : t_hour) * 64);      // half `odd` of the even lane's row
    const U256 h1 = ldg256_nc(base + (int64_t)(odd ? t_hour : t_p) * 64);      // half `odd` of the odd lane's row
    float4* r0 = reinterpret_cast<float4*>(sm + (lane - odd) * 40);            // row 2p; row 2p + 1 follows 10 groups later
    const int g = odd ? 8 : 6;                                                 // Pot_Reward 8..11 | 0..3
    r0[g] = f4_lo(h0.a, h0.b);
    r0[10 + g] = f4_lo(h1.a, h1.b);
    if (!odd) {                                                                // Pot_Reward 4..7
        r0[7] = f4_lo(h0.c, h0.d);
        r0[17] = f4_lo(h1.c, h1.d);
    }
    const unsigned long long c_p = __shfl_xor_sync(0xffffffffu, h0.c, 1), d_p = __shfl_xor_sync(0xffffffffu, h0.d, 1);
    const unsigned long long c_own = odd ? h1.c : c_p, d_own = odd ? h1.d : d_p;
    pr12 = __uint_as_float((uint32_t)c_own);
    const int bits = (int)(c_own >> 32);
    auto pf = [bits](int a) { return (float)((bits << (30 - 2 * a)) >> 30); };
    float4* r4 = reinterpret_cast<float4*>(sm + lane * 40);
    r4[3] = make_float4(pf(1), pf(2), pf(3), pf(4));
    r4[4] = make_float4(pf(5), pf(6), pf(7), pf(8));
    r4[5] = make_float4(pf(9), pf(10), pf(11), pf(12));
    pf0 = pf(0);
    __syncwarp();          // (a lane that ends its episode re-stages its own row later: order it behind the partner's stores)
    return __longlong_as_double((long long)d_own);
}

template <int NV>
__device__ __forceinline__ void load_hour_row(const DevParams& P, int t_hour, float4 (&hrow)[NV]) {
    const float4* src = P.hour_tab + (int64_t)t_hour * NV;
    if (NV % 2 == 0) {                      // rows are multiples of 32 B: 256-bit gathers
#pragma unroll
        for (int v = 0; v < NV; v += 2) {
            const U256 r = ldg256_nc(src + v);
            hrow[v] = f4_lo(r.a, r.b);
            hrow[v + 1 < NV ? v + 1 : v] = f4_lo(r.c, r.d);
        }
    } else {
#pragma unroll
        for (int v = 0; v < NV; ++v) hrow[v] = __ldg(src + v);
    }
}

__device__ __forceinline__ DayRow load_day_row(const DevParams& P, int t_day) {
    const U256 r = ldg256_nc(P.day_tab + t_day);
    DayRow d;
    d.gas = __longlong_as_double((long long)r.a); d.eua = __longlong_as_double((long long)r.b);
    d.gas_n0 = __uint_as_float((uint32_t)r.c); d.gas_n1 = __uint_as_float((uint32_t)(r.c >> 32));
    d.eua_n0 = __uint_as_float((uint32_t)r.d); d.eua_n1 = __uint_as_float((uint32_t)(r.d >> 32));
    return d;
}

template <int NV>
__device__ __forceinline__ double hour_row_el(const float4 (&hrow)[NV]) {
    return __hiloint2double(__float_as_int(hrow[NV - 1].w), __float_as_int(hrow[NV - 1].z));
}

__device__ __forceinline__ void clamp_market_index(const DevParams& P, int& t_hour, int& t_day) {
    if (t_hour >= P.n_hours || t_day >= P.n_days || t_hour < 0 || t_day < 0) {
        atomicOr(P.err, PTG_EBIT_RANGE);       // the reference raises IndexError here
        t_hour = max(0, min(t_hour, P.n_hours - 1));
        t_day = max(0, min(t_day, P.n_days - 1));
    }
}

// _get_info, :251-278, feature-major fp64 [24][n_envs]
struct InfoKey {
    int k, t_hour, t_day, ent, state_change;   // ent < 0: reset (no reward constituents yet)
    uint32_t meta;
    double rew, cum_rew;
};
template <int NV>
__device__ __noinline__ void write_info(const DevParams& P, double* info, int64_t e, InfoKey s) {
    const int64_t n = P.n_envs;
    double* f = info + e;
    const Meta m = meta_unpack(s.meta);
    const double el = *reinterpret_cast<const double*>(reinterpret_cast<const float*>(P.hour_tab + (int64_t)s.t_hour * NV) + 4 * NV - 2);
    const DayRow day = P.day_tab[s.t_day];
    double flow[5], t_cat;
    RewardParts r = {};
    if (s.ent >= 0) {          // constituents in the reference's operation order, from the fp64 window means
        const StepMeans sm = P.mean_tab[s.ent];
#pragma unroll
        for (int q = 0; q < 5; ++q) flow[q] = sm.mean[q];
        t_cat = sm.t_end;
        reward_parts(P.rc, flow, el, day.gas, day.eua, s.state_change, r);
    } else {
#pragma unroll
        for (int q = 0; q < 5; ++q) flow[q] = P.reset_flow[q];
        t_cat = 16.0;
    }
    f[0 * n] = s.k;
    f[1 * n] = el;
    f[2 * n] = day.gas;
    f[3 * n] = day.eua;
    f[4 * n] = m.state;
    f[5 * n] = m.cur_action;
    f[6 * n] = m.hot_cold;
    f[7 * n] = t_cat;
    f[8 * n] = flow[0]; f[9 * n] = flow[1]; f[10 * n] = flow[3]; f[11 * n] = flow[4];
    f[12 * n] = r.ch4_rev; f[13 * n] = r.steam_rev; f[14 * n] = r.o2_rev; f[15 * n] = r.eua_rev; f[16 * n] = r.chp_rev;
    f[17 * n] = -r.heat_cost; f[18 * n] = -r.ely_cost; f[19 * n] = -r.water_cost;
    f[20 * n] = s.rew;         // "reward [ct]" is the value step() returned
    f[21 * n] = s.cum_rew;
    f[22 * n] = P.pot0[s.t_hour];
    f[23 * n] = P.pf0[s.t_hour];
}

// _initialize_op_rew (:105-138) + episode offsets; returns the reset (core, tinfo, ep)
__device__ __forceinline__ void env_reset_state(const DevParams& P, int64_t e, int32_t m_count, int4& core,
                                                int32_t& tinfo, int2& ep, Meta& m) {
    int ep_h, ep_d;
    episode_offsets(P, e, m_count, ep_h, ep_d);
    ep = make_int2(ep_h, ep_d);
    m.state = PTG_COOLDOWN; m.hot_cold = 0; m.standby_ds = PTG_DS_STANDBY_DOWN; m.startup_ds = PTG_DS_STARTUP_COLD;
    m.part_ds = PTG_DS_OP1_START_P; m.full_ds = PTG_DS_OP2_START_F;
    // current_action is NOT touched by reset() (only by __init__, :143)
    core = make_int4(P.reset_i, 0, 0, (int)meta_pack(m));
    tinfo = (int32_t)((uint32_t)P.reset_tinfo | ((uint32_t)tinfo & ~PTG_TI_LOW_MASK));    // (in/out: draw counter kept)
}

// ------------------------------------------------------------------------------------------------------------
// constructor / reset
// ------------------------------------------------------------------------------------------------------------
__global__ void k_construct(const __grid_constant__ DevParams P) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= P.n_envs) return;
    const int64_t gid = P.env_id_offset + e;
    // The reference's generator is unseeded until reset(seed=...) (gymnasium); we start every env from
    // SeedSequence(global id) so that an unseeded run is still reproducible.
    Pcg64 g = pcg64_from_seed((uint64_t)gid);
    RngRec rr = {};
    rr.s_hi = g.s_hi; rr.s_lo = g.s_lo; rr.i_hi = g.i_hi; rr.i_lo = g.i_lo;
    P.rng[e] = rr;
    P.draws_total[e] = 0;
    if (P.schedule_mode == PTG_SCHED_SUBPROC) {
        // :43-44 draws ep_index = integers(0, n_eps_loops) from an UNSEEDED generator per worker process (not
        // reproducible in the reference); here: a fixed stream per global env id, same range
        Pcg64 h = pcg64_from_seed(0x5eed0000ull + (uint64_t)gid);
        const int range = max(1, min(P.n_eps_loops, P.n_eps_ind > 0 ? P.n_eps_ind : 1));
        P.ep_start[e] = (int32_t)(pcg64_next64(h) % (uint64_t)range);
    } else {
        P.ep_start[e] = 0;
    }
    Meta m;
    m.cur_action = PTG_COOLDOWN;     // :143
    int4 core; int32_t tinfo = 0; int2 ep;
    env_reset_state(P, e, 0, core, tinfo, ep, m);
    P.core[e] = core; P.tinfo[e] = tinfo; P.ep[e] = ep;
    P.ep_ret[e] = 0.0; P.ep_count[e] = 0;
    if (P.has_penalty) P.nchg[e] = 0;
    P.fin_cnt[e] = 0; P.fin_ret_sum[e] = 0.0; P.fin_ret_sq[e] = 0.0; P.fin_len_sum[e] = 0.0;
    P.fin_min[e] = INFINITY; P.fin_max[e] = -INFINITY;
}

template <int NV, bool MOD, bool FLAT = false>
__global__ void __launch_bounds__(PTG_BLOCK) k_reset(const __grid_constant__ DevParams P, const int64_t* seeds,
                                                     const uint8_t* mask, const __grid_constant__ PtgIO io) {
    __shared__ __align__(128) float stage[PTG_BLOCK / 32][2 * PTG_STAGE_FLOATS(NV)];
    const int e = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int warp_env0 = e - lane;
    const int nvalid = min(32, (int)P.n_envs - warp_env0);
    const bool active = e < (int)P.n_envs;
    const bool doit = active && (mask == nullptr || mask[e]);
    // masked reset: only the selected envs write; the window blocks go through the scalar path then
    float4 hrow[NV];
    DayRow day = {};
    ObsRegs o = {};
#pragma unroll
    for (int v = 0; v < NV; ++v) hrow[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    int t_hour = 0, t_day = 0;
    Meta m = {};
    if (doit) {
        int32_t tinfo = P.tinfo[e];                       // (its draw-counter bits survive a plain reset)
        if (seeds != nullptr && seeds[e] >= 0) {          // gymnasium Env.reset(seed=...): a fresh generator
            Pcg64 g = pcg64_from_seed((uint64_t)seeds[e]);
            RngRec rr = {};
            rr.s_hi = g.s_hi; rr.s_lo = g.s_lo; rr.i_hi = g.i_hi; rr.i_lo = g.i_lo;
            P.rng[e] = rr;
            P.draws_total[e] = 0;
            tinfo = 0;
        }
        m = meta_unpack((uint32_t)P.core[e].w);
        const int32_t mc = P.ep_count[e] + 1;
        int4 core; int2 ep;
        env_reset_state(P, e, mc, core, tinfo, ep, m);
        P.core[e] = core; P.tinfo[e] = tinfo; P.ep[e] = ep; P.ep_count[e] = mc; P.ep_ret[e] = 0.0;
        if (P.has_penalty) P.nchg[e] = 0;
        t_hour = ep.x; t_day = ep.y;
        clamp_market_index(P, t_hour, t_day);
        load_hour_row<NV>(P, t_hour, hrow);
        day = load_day_row(P, t_day);
        o.status = PTG_COOLDOWN;
#pragma unroll
        for (int q = 0; q < 6; ++q) o.norm[q] = P.reset_norm[q];
        o.sin_h = 0.0f; o.cos_h = 1.0f;                   // math.sin(0), math.cos(0), :102-103
    }
    if (io.obs != nullptr) {
        if (FLAT) { if (doit) emit_flat_row<MOD>(P, io.obs, e, ObsKey{-1, t_hour, t_day, PTG_COOLDOWN, 0}); }
        else if (mask == nullptr) emit_obs<NV, MOD>(P, io.obs, stage[wid], e, active, lane, warp_env0, nvalid, hrow, day, o);
        else if (doit) emit_obs_scalar<NV, MOD>(P, io.obs, e, ObsKey{-1, t_hour, t_day, PTG_COOLDOWN, 0});
    }
    if (elect_one()) tma_store_wait_read();
    if (doit && io.info != nullptr) {
        write_info<NV>(P, io.info, e, InfoKey{0, t_hour, t_day, -1, 0, meta_pack(m), 0.0, 0.0});
    }
}

// ------------------------------------------------------------------------------------------------------------
// step
// ------------------------------------------------------------------------------------------------------------
// End of an episode (rare): record the terminal observation and the Monitor record, fold the episode into the
// finished-episode accumulators and write the reset state (next entry of the episode schedule) to global memory.
template <int NV, bool MOD, bool FLAT = false>
__device__ __noinline__ void finish_episode(const DevParams& P, const PtgIO& io, int64_t e, bool record, int k,
                                            double ep_ret, int cur_action, ObsKey key, uint32_t draws_ep) {
    if (record) {
        if (io.terminal_obs != nullptr) {
            if (FLAT) emit_flat_row<MOD>(P, io.terminal_obs, e, key);
            else emit_obs_scalar<NV, MOD>(P, io.terminal_obs, e, key);
        }
        if (io.episode_return != nullptr) io.episode_return[e] = ep_ret;
        if (io.episode_length != nullptr) io.episode_length[e] = k;
    }
    P.fin_cnt[e] += 1;
    P.fin_ret_sum[e] += ep_ret;
    P.fin_ret_sq[e] += ep_ret * ep_ret;
    P.fin_len_sum[e] += (double)k;
    P.fin_min[e] = fmin(P.fin_min[e], ep_ret);
    P.fin_max[e] = fmax(P.fin_max[e], ep_ret);
    if (!P.auto_reset) return;           // Gymnasium single-env semantics: the caller resets (PtgConfig.auto_reset = 0)
    const int32_t mc = P.ep_count[e] + 1;
    P.ep_count[e] = mc;
    Meta m;
    m.cur_action = cur_action;
    int4 core; int32_t tinfo = (int32_t)(draws_ep << PTG_TI_DRAW_SHIFT); int2 ep;
    env_reset_state(P, e, mc, core, tinfo, ep, m);
    P.core[e] = core; P.tinfo[e] = tinfo; P.ep[e] = ep; P.ep_ret[e] = 0.0;
    if (P.has_penalty) P.nchg[e] = 0;
}

#ifndef PTG_PDL
#define PTG_PDL 1
#endif
#ifndef PTG_PAIR_GATHER
#define PTG_PAIR_GATHER 1        // step-table gather: lane pairs split their two entries (half the L1 wavefronts, +29 instructions)
#endif
#ifndef PTG_PERSIST_ALL
#define PTG_PERSIST_ALL 1        // single steps run persistent CTAs in both layouts (0: key-major with one tile per CTA, round 1)
#endif
#ifndef PTG_RNG_PREFETCH
#define PTG_RNG_PREFETCH 0       // 1: L1 prefetch of a drawing lane's RNG record as soon as (action, state) are known (round 1;
#endif                           //    it costs the L1TEX data pipe as many wavefronts as the load it hides: 53.2 vs 52.3 us without)
#ifndef PTG_STEP_MIN_BLOCKS
#define PTG_STEP_MIN_BLOCKS 4        // CTAs of 256 threads per SM the step kernel is compiled for (<= 64 registers)
#endif

// Per-env plant state (core, tinfo, ep, ep_ret: 36 B read + 28 B written per env-step).  PTG_STATE_L2 = 1 keeps it in
// L2 between steps (evict_last on loads and stores; 36 MB at 1 M envs next to 32 MB of RNG records and 16 MB of tables
// in the 126 MB L2, while observations / rewards / actions stream with evict_first); 0 = streamed (round-1 behaviour).
#ifndef PTG_STATE_L2
#define PTG_STATE_L2 1
#endif
#ifndef PTG_PREFETCH_STATE
#define PTG_PREFETCH_STATE 0     // 1: prefetch.global.L2 of the plant state of the CTA one scheduling wave later (the
#endif                           //    round-1 arrangement for streamed state; pointless once the state lives in L2, and
                                 //    its address arithmetic cost the roll-out kernel two spilled registers: 34.8 -> 33.4 us)
#if PTG_STATE_L2 && PTG_L2_HINTS
#define PTG_LD_STATE(p) ld_keep(p)
#define PTG_ST_STATE(p, v) st_keep(p, v)
#else
#define PTG_LD_STATE(p) (*(p))
#define PTG_ST_STATE(p, v) st_stream(p, v)
#endif

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ long long load_action_raw(const void* __restrict__ actions, int adtype, int64_t idx) {
    // streamed once per step (ld.global.cs: evict-first), so the action tensor does not displace L2-resident state
    if (adtype == PTG_ACT_I64) return __ldcs((const long long*)actions + idx);
    if (adtype == PTG_ACT_I32) return __ldcs((const int*)actions + idx);
    if (adtype == PTG_ACT_U8) return __ldcs((const unsigned char*)actions + idx);
    return (long long)__float_as_int(__ldcs((const float*)actions + idx));      // F32: raw bits
}

// Decode the action of an env (discrete id, or continuous Box(-1,1) -> 5 bins, :346-355) from load_action_raw.
__device__ __forceinline__ int decode_action_raw(const DevParams& P, long long raw, int adtype, int prev_action) {
    if (!P.continuous) {
        long long a = adtype == PTG_ACT_F32 ? (long long)__int_as_float((int)raw) : raw;
        if (a < 0 || a > 4) { atomicOr(P.err, PTG_EBIT_ACTION); a = PTG_COOLDOWN; }
        return (int)a;
    }
    const double a = (double)__int_as_float((int)raw);
    int act = prev_action;                        // a >= 1.0: no interval matches, previous action is kept
#pragma unroll
    for (int ival = 5; ival >= 0; --ival)
        if (P.prob_thre[ival] > a) act = (ival + 4) % 5;   // first matching ival wins (descending scan)
    return act;
}

// Step-table gather (one 64 B entry per env).  L1TEX cost of a gather is per (instruction, 128 B line): in a full
// warp a lane PAIR splits its two entries so that each LDG.256 touches 16 lines instead of 32 (even lane: first
// halves, odd lane: second halves), then the halves are exchanged with four 64-bit shuffles.
__device__ __forceinline__ void gather_step_entry(const DevParams& P, int ent, int lane, int nvalid, U256& qc, U256& qn) {
    if (nvalid == 32 && PTG_PAIR_GATHER) {
        const int ent_p = __shfl_xor_sync(0xffffffffu, ent, 1);
        const bool odd = lane & 1;
        const char* base = reinterpret_cast<const char*>(P.step_tab) + (odd ? 32 : 0);
        const U256 a = ldg256_nc(base + (int64_t)(odd ? ent_p : ent) * 64);     // the even lane's entry
        const U256 b = ldg256_nc(base + (int64_t)(odd ? ent : ent_p) * 64);     // the odd lane's entry
        U256 snd = odd ? a : b, rcv;
        rcv.a = __shfl_xor_sync(0xffffffffu, snd.a, 1); rcv.b = __shfl_xor_sync(0xffffffffu, snd.b, 1);
        rcv.c = __shfl_xor_sync(0xffffffffu, snd.c, 1); rcv.d = __shfl_xor_sync(0xffffffffu, snd.d, 1);
        qc = odd ? rcv : a;
        qn = odd ? b : rcv;
    } else {
        qc = ldg256_nc(P.step_tab + ent);
        qn = ldg256_nc(reinterpret_cast<const char*>(P.step_tab + ent) + 32);
    }
}

// One env step of one thread.  Ordering (each group only depends on the ones above it):
//   1. argmin-LUT gather + prefetch of the RNG line      <- only needs (state, action)
//   2. clock row -> market rows -> stage the two observation windows in shared memory (independent of the
//      plant transition: covers the latency of 1.)
//   3. plant transition (may draw noise) -> step-table entry gather -> reward, scalars
//   4. bulk stores
template <int NV, bool MOD, bool EVAL, int PAC>
__device__ __forceinline__ void step_one(const DevParams& P, const PtgIO& io, long long action_raw, int adtype,
                                         bool single, int e, bool active, int lane, int warp_env0,
                                         int nvalid, float* sm, const uint64_t* zig_kiwi, const uint16_t* plan_lut,
                                         float* __restrict__ obs_out,
                                         float* __restrict__ rew_out, uint8_t* __restrict__ done_out, int& i, int& j,
                                         int& k, uint32_t& meta, int32_t& tinfo, int2& ep, double& ep_ret) {
    float4 hrow[NV];
    DayRow day = {};
    ObsRegs o;
    float reward = 0.f;
    int t_hour_out = 0;
    // termination only depends on the step counter (:508-511, k before the increment): known up front, so the
    // common case (no env of the warp ends its episode) can hand its window tiles to the TMA early
    const int done = active && (k == P.eps_sim_steps - 6);
    const bool any_done = __any_sync(0xffffffffu, done);
    bool win_moved = false;
    if (active) {
        // (1) what the transition will need from memory
        const int action = decode_action_raw(P, action_raw, adtype, (meta >> 4) & 7);
        const int prev_state = meta & 7;
        const Plan plan = plan_transition(action, meta, tinfo & 7, plan_lut);
        uint32_t draws_ep = ti_draws(tinfo);
        int lut_val = 0;
        if (plan.col >= 0) {
            int vid = ti_id(tinfo);
            PTG_CHECK_INDEX(P, vid, P.n_vals, 4);
            lut_val = ldg32_nc_keep(P.argmin_lut + vid * PTG_N_ARGMIN + plan.col);
        }
        uint32_t chain_c = 0;
        if (PTG_CHAIN_EARLY && plan.kind >= PTG_KIND_PARTIAL) chain_c = chain_lookup(P, plan, meta, i, j);
        const bool draws = plan.kind == PTG_KIND_DRAW && P.noise_mode != PTG_NOISE_OFF;
        if (PTG_RNG_PREFETCH && draws) prefetch_l1(P.rng + e);
        // (2) clock of step k+1 (:442-445, integer form of floor(clock_hours), floor(clock_days)) -> market rows of
        //     the NEW hour/day (:446-447) -> observation windows; sin/cos of the clock come from the clock table
        const unsigned sec = (unsigned)(k + 1) * (unsigned)P.sim_step;
        int t_hour = ep.x + (int)(sec / 3600u), t_day = ep.y + (int)(sec / 86400u);
        // the market-window blocks of the observation only move when the clock crosses an hour (or the episode ends)
        win_moved = done || (sec / 3600u) != ((sec - (unsigned)P.sim_step) / 3600u);
        clamp_market_index(P, t_hour, t_day);
        int k1 = k + 1;
        PTG_CHECK_INDEX(P, k1, P.eps_sim_steps + 1, 5);
        // quad rows carry the day's prices: usable by a full warp whose lanes all have t_day == t_hour / 24
        const bool quad = PTG_HOUR_QUAD && MOD && NV == 4 && PAC == 13 && nvalid == 32 && P.hourq_tab != nullptr &&
                          __all_sync(0xffffffffu, (unsigned)(t_hour - 24 * t_day) < 24u);
        if (!quad) day = load_day_row(P, t_day);
        const float2 sc2 = __ldg(reinterpret_cast<const float2*>(P.clock_tab + k1));
        double el, gas = day.gas, eua = day.eua;
        if (quad) {
            el = stage_windows_quad(P, sm, lane, t_hour, gas, eua);
        } else if (hour_interleaved(NV, false) && PAC == 13 && nvalid == 32) {
            el = stage_windows_paired<MOD>(P, sm, lane, t_hour);
        } else {
            load_hour_row<NV>(P, t_hour, hrow);
            stage_windows<NV, MOD, PAC>(P, sm, lane, hrow, nvalid == 32);
            el = hour_row_el<NV>(hrow);               // from here on the hour row is dead (registers!)
        }
        // the window tiles go to the TMA before the transition: the fence in front of a bulk store waits for the
        // thread's outstanding accesses, so it must not sit behind the RNG / step-table requests
        if (!any_done) issue_window_stores<NV, MOD, PAC>(P, obs_out, sm, lane, warp_env0, nvalid);
        // (3) plant transition (requests the RNG record when it draws) -> step-table entry (2 x 32 B sectors)
        const int ent = apply_transition(P, e, plan, i, j, meta, lut_val, zig_kiwi, draws_ep, chain_c);
        const int state_change = (prev_state != (int)(meta & 7));
        U256 qc, qn;      // qc = {c_gas, c_eua, c_el, c_0}, qn = {norm[6], tinfo, pad}
        gather_step_entry(P, ent, lane, nvalid, qc, qn);
        // reward (:280-334 in price-linear form) with the prices of the new hour/day (:463-468)
        const double c_gas = __longlong_as_double((long long)qc.a), c_eua = __longlong_as_double((long long)qc.b);
        const double c_el = __longlong_as_double((long long)qc.c), c_0 = __longlong_as_double((long long)qc.d);
        double rew = __fma_rn(c_gas, gas, __fma_rn(c_eua, eua, __fma_rn(-c_el, el, c_0)));
        if (state_change) rew -= P.penalty;
        ep_ret += rew;
        reward = (float)rew;
        o.norm[0] = __uint_as_float((uint32_t)qn.a); o.norm[1] = __uint_as_float((uint32_t)(qn.a >> 32));
        o.norm[2] = __uint_as_float((uint32_t)qn.b); o.norm[3] = __uint_as_float((uint32_t)(qn.b >> 32));
        o.norm[4] = __uint_as_float((uint32_t)qn.c); o.norm[5] = __uint_as_float((uint32_t)(qn.c >> 32));
        tinfo = (int32_t)(((uint32_t)qn.d & PTG_TI_LOW_MASK) | (draws_ep << PTG_TI_DRAW_SHIFT));
        o.status = meta & 7;
        o.sin_h = sc2.x; o.cos_h = sc2.y;
        uint32_t nchg = 0;
        if (P.has_penalty) { nchg = P.nchg[e] + (uint32_t)state_change; P.nchg[e] = nchg; }
        if (EVAL && io.info != nullptr)
            write_info<NV>(P, io.info, e, InfoKey{k, t_hour, t_day, ent, state_change, meta, rew,
                                                  ep_ret + (double)nchg * P.penalty});
        k += 1;
        if (done) finish_episode<NV, MOD>(P, io, e, single, k, ep_ret, (meta >> 4) & 7,
                                          ObsKey{ent, t_hour, t_day, (int)(meta & 7), k}, draws_ep);
        if (done && P.auto_reset) {                   // SB3 auto-reset (DummyVecEnv.step_wait), out of line
            const int4 core = P.core[e];              // the reset state written by finish_episode
            tinfo = P.tinfo[e]; ep = P.ep[e];
            i = core.x; j = core.y; k = core.z; meta = (uint32_t)core.w; ep_ret = 0.0;
            t_hour = ep.x; t_day = ep.y;
            clamp_market_index(P, t_hour, t_day);
            day = load_day_row(P, t_day);
            o.status = PTG_COOLDOWN;
#pragma unroll
            for (int q = 0; q < 6; ++q) o.norm[q] = P.reset_norm[q];
            o.sin_h = 0.0f; o.cos_h = 1.0f;
        }
        t_hour_out = t_hour;
    } else {
#pragma unroll
        for (int v = 0; v < NV; ++v) hrow[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        day = DayRow{};
        o = ObsRegs{};
        stage_windows<NV, MOD, PAC>(P, sm, lane, hrow, nvalid == 32);
        if (!any_done) issue_window_stores<NV, MOD, PAC>(P, obs_out, sm, lane, warp_env0, nvalid);
    }
    // episode ends are rare: only then is the warp's window tile staged again, from re-read hour rows (the reset
    // observation of the done lanes, the unchanged rows of the others), and stored after the fact
    if (any_done) {
        float4 h2[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) h2[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) load_hour_row<NV>(P, t_hour_out, h2);
        stage_windows<NV, MOD, PAC>(P, sm, lane, h2, nvalid == 32);
        issue_window_stores<NV, MOD, PAC>(P, obs_out, sm, lane, warp_env0, nvalid);
    }
    if (active) {
        store_obs_scalars<MOD>(P, obs_out, e, day, o);
        st_stream(rew_out + e, reward);
        st_stream(done_out + e, (uint8_t)done);
        if (single && io.status_u8 != nullptr) st_stream(io.status_u8 + e, (uint8_t)o.status);
    }
    if (single && io.windows_changed != nullptr && __any_sync(0xffffffffu, win_moved) && lane == 0)
        *io.windows_changed = P.step_serial;        // (same value from every warp: plain store, no atomic needed)
}

// One env step with the flat observation layout: the same transition / reward path as step_one, but the whole
// observation of a warp -- 32 rows of F floats, contiguous in the output -- is assembled in shared memory and leaves
// the SM as ONE bulk store (5 120 B for mod) at the end of the step; no scalar observation stores at all.
template <bool MOD>
__device__ __forceinline__ void step_one_flat(const DevParams& P, const PtgIO& io, long long action_raw, int adtype,
                                              bool single, int e, bool active, int lane, int warp_env0, int nvalid,
                                              float* sm, const uint64_t* zig_kiwi, const uint16_t* plan_lut,
                                              float* __restrict__ obs_out,
                                              float* __restrict__ rew_out, uint8_t* __restrict__ done_out, int& i,
                                              int& j, int& k, uint32_t& meta, int32_t& tinfo, int2& ep, double& ep_ret) {
    constexpr int F = PTG_FLAT_F(MOD);
    float4 hrow[4];
    DayRow day;
    ObsRegs o;
    float reward = 0.f, pf0 = 0.f, pr12 = 0.f;
    const int done = active && (k == P.eps_sim_steps - 6);
    float* row = sm + lane * F;
    if (elect_one()) tma_store_wait_read();       // the previous step's bulk store (rollout kernel) is done with the tile
    __syncwarp();
    if (active) {
        const int action = decode_action_raw(P, action_raw, adtype, (meta >> 4) & 7);
        const int prev_state = meta & 7;
        const Plan plan = plan_transition(action, meta, tinfo & 7, plan_lut);
        uint32_t draws_ep = ti_draws(tinfo);
        int lut_val = 0;
        if (plan.col >= 0) {
            int vid = ti_id(tinfo);
            PTG_CHECK_INDEX(P, vid, P.n_vals, 4);
            lut_val = ldg32_nc_keep(P.argmin_lut + vid * PTG_N_ARGMIN + plan.col);
        }
        uint32_t chain_c = 0;
        if (PTG_CHAIN_EARLY && plan.kind >= PTG_KIND_PARTIAL) chain_c = chain_lookup(P, plan, meta, i, j);
        if (PTG_RNG_PREFETCH && plan.kind == PTG_KIND_DRAW && P.noise_mode != PTG_NOISE_OFF) prefetch_l1(P.rng + e);
        const unsigned sec = (unsigned)(k + 1) * (unsigned)P.sim_step;
        int t_hour = ep.x + (int)(sec / 3600u), t_day = ep.y + (int)(sec / 86400u);
        clamp_market_index(P, t_hour, t_day);
        day = load_day_row(P, t_day);
        int k1 = k + 1;
        PTG_CHECK_INDEX(P, k1, P.eps_sim_steps + 1, 5);
        const float2 sc2 = __ldg(reinterpret_cast<const float2*>(P.clock_tab + k1));
        double el;
        if (MOD && PTG_FLAT_PAIRED && nvalid == 32) {
            el = stage_flat_early_paired(P, sm, lane, t_hour, pf0, pr12);
        } else {
            load_hour_row<4>(P, t_hour, hrow);
            stage_flat_early<MOD>(row, hrow, day, pf0, pr12);
            el = hour_row_el<4>(hrow);
        }
        const int ent = apply_transition(P, e, plan, i, j, meta, lut_val, zig_kiwi, draws_ep, chain_c);
        const int state_change = (prev_state != (int)(meta & 7));
        U256 qc, qn;
        gather_step_entry(P, ent, lane, nvalid, qc, qn);
        const double c_gas = __longlong_as_double((long long)qc.a), c_eua = __longlong_as_double((long long)qc.b);
        const double c_el = __longlong_as_double((long long)qc.c), c_0 = __longlong_as_double((long long)qc.d);
        double rew = __fma_rn(c_gas, day.gas, __fma_rn(c_eua, day.eua, __fma_rn(-c_el, el, c_0)));
        if (state_change) rew -= P.penalty;
        ep_ret += rew;
        reward = (float)rew;
        o.norm[0] = __uint_as_float((uint32_t)qn.a); o.norm[1] = __uint_as_float((uint32_t)(qn.a >> 32));
        o.norm[2] = __uint_as_float((uint32_t)qn.b); o.norm[3] = __uint_as_float((uint32_t)(qn.b >> 32));
        o.norm[4] = __uint_as_float((uint32_t)qn.c); o.norm[5] = __uint_as_float((uint32_t)(qn.c >> 32));
        tinfo = (int32_t)(((uint32_t)qn.d & PTG_TI_LOW_MASK) | (draws_ep << PTG_TI_DRAW_SHIFT));
        o.status = meta & 7;
        o.sin_h = sc2.x; o.cos_h = sc2.y;
        if (P.has_penalty) P.nchg[e] += (uint32_t)state_change;
        k += 1;
        if (done) finish_episode<4, MOD, true>(P, io, e, single, k, ep_ret, (meta >> 4) & 7,
                                               ObsKey{ent, t_hour, t_day, (int)(meta & 7), k}, draws_ep);
        if (done && P.auto_reset) {                   // SB3 auto-reset: the returned row is the reset observation
            const int4 core = P.core[e];
            tinfo = P.tinfo[e]; ep = P.ep[e];
            i = core.x; j = core.y; k = core.z; meta = (uint32_t)core.w; ep_ret = 0.0;
            t_hour = ep.x; t_day = ep.y;
            clamp_market_index(P, t_hour, t_day);
            load_hour_row<4>(P, t_hour, hrow);
            day = load_day_row(P, t_day);
            stage_flat_early<MOD>(row, hrow, day, pf0, pr12);
            o.status = PTG_COOLDOWN;
#pragma unroll
            for (int q = 0; q < 6; ++q) o.norm[q] = P.reset_norm[q];
            o.sin_h = 0.0f; o.cos_h = 1.0f;
        }
        stage_flat_late<MOD>(row, o, pf0, pr12);
    }
    float* g = obs_out + (int64_t)warp_env0 * F;
    const uint32_t bytes = (uint32_t)(nvalid * F) * 4u;
    if ((bytes & 15u) == 0) {
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) { tma_store_1d(g, sm, bytes); tma_store_commit(); }
    } else {                                          // ragged tail of the raw design: plain stores
        __syncwarp();
        for (int idx = lane; idx < nvalid * F; idx += 32) g[idx] = sm[idx];
    }
    if (active) {
        st_stream(rew_out + e, reward);
        st_stream(done_out + e, (uint8_t)done);
        if (single && io.status_u8 != nullptr) st_stream(io.status_u8 + e, (uint8_t)o.status);
    }
}

// VecEnv.step_wait(): MANY = false -> exactly one step (ptg_step); MANY = true -> T steps with the plant state
// kept in registers between steps (ptg_step_many).  EVAL = the 24-field info of train_or_eval == "eval".
template <int NV, bool MOD, bool MANY, bool EVAL, int PAC, bool FLAT = false>
__global__ void __launch_bounds__(PTG_BLOCK, PTG_STEP_MIN_BLOCKS)
k_step(const __grid_constant__ DevParams P, const void* __restrict__ actions, int adtype,
       const __grid_constant__ PtgIO io, int T) {
    static_assert(!FLAT || (NV == 4 && PAC == 13 && !EVAL), "the flat layout is built for price_ahead == 13");
    __shared__ __align__(128) float stage[PTG_BLOCK / 32][FLAT ? 32 * PTG_FLAT_F(MOD) : 2 * PTG_STAGE_FLOATS(NV)];
    __shared__ __align__(16) uint64_t zig_kiwi[2 * 256];   // {ki, wi} of numpy's ziggurat: 4 KB, one LDS.128 per draw
    __shared__ __align__(4) uint16_t plan_lut[PTG_PLAN_LUT_SIZE];   // plan_pack() of every (action, state, T flags, hot_cold)
    static_assert(256 % PTG_BLOCK == 0, "the CTA stages the 256 ziggurat layers in 256 / PTG_BLOCK rounds");
    const int n_envs = (int)P.n_envs;                   // < 2^26 (checked by ptg_create): 32-bit index arithmetic
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool use_zig = P.noise_mode == PTG_NOISE_NUMPY;
#if PTG_PDL
    // Programmatic dependent launch: the next launch in the stream may become resident as soon as every CTA of this
    // one has started, so its prologue (ziggurat table staging below) overlaps this kernel's tail wave ...
    asm volatile("griddepcontrol.launch_dependents;");
#endif
    for (int q = threadIdx.x; q < PTG_PLAN_LUT_SIZE / 2; q += PTG_BLOCK)
        reinterpret_cast<uint32_t*>(plan_lut)[q] = __ldg(reinterpret_cast<const uint32_t*>(P.plan_lut) + q);
    if (use_zig) {
#pragma unroll
        for (int q = threadIdx.x; q < 256; q += PTG_BLOCK) {
            zig_kiwi[2 * q] = __ldg(P.zig.ki + q);
            zig_kiwi[2 * q + 1] = (uint64_t)__double_as_longlong(__ldg(P.zig.wi + q));
        }
    }
#if PTG_PDL
    // ... and nothing the previous kernel may still be writing (env state, actions) is touched before it completed
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
    // Single steps run PERSISTENT CTAs (one scheduling wave of them, tile b, b + gridDim, ...): a warp that finishes a
    // tile starts its next one without waiting for the seven other warps of its CTA to drain and for a new CTA to be
    // scheduled, and the ziggurat table is staged once per CTA instead of once per tile.  Flat layout: 63.7 -> 58.1 us
    // (round 1: its end-of-step bulk store otherwise holds the CTA's slot while the TMA reads the tile).  Key-major
    // layout: round 1 measured the opposite (56.5 vs 54 us, with the L2 state prefetch); with the plant state resident
    // in L2 and the prefetch gone it is 56.5 vs 58.3 us uniform, 41.8 vs 42.7 us sticky (round 2, same box).
    constexpr bool PERSIST = (FLAT || PTG_PERSIST_ALL) && !MANY;   // (the roll-out kernel's stores overlap its next step anyway)
    if (PERSIST) __syncthreads();                       // staged tables visible to every warp of the CTA
    for (int tile = blockIdx.x; tile * PTG_BLOCK < n_envs; tile += gridDim.x) {
    const int e = tile * PTG_BLOCK + (int)threadIdx.x;
    const int warp_env0 = e - lane;
    const bool warp_in_range = warp_env0 < n_envs;
    const int nvalid = min(32, n_envs - warp_env0);
    const bool active = e < n_envs;
    const int le = active ? e : n_envs - 1;             // tail lanes shadow the last env (no stores)

    const int4 core = PTG_LD_STATE(P.core + le);
    int32_t tinfo = PTG_LD_STATE(P.tinfo + le);
    int2 ep = PTG_LD_STATE(P.ep + le);
    double ep_ret = PTG_LD_STATE(P.ep_ret + le);
    long long action_raw = load_action_raw(actions, adtype, le);
    if (!PERSIST) __syncthreads();                      // (one tile per CTA: the barrier sits behind the state loads)
    if (!warp_in_range) continue;                       // whole warp out of range
    int i = core.x, j = core.y, k = core.z;
    uint32_t meta = (uint32_t)core.w;
#if PTG_PREFETCH_STATE
    {   // pull the plant state of the CTA that will run one scheduling wave later into L2 (fire and forget)
        const int pe = e + P.prefetch_distance;
        if (pe < n_envs) {
            prefetch_l2(P.core + pe);
            if ((lane & 3) == 0) prefetch_l2(P.ep_ret + pe);
            if ((lane & 3) == 1) prefetch_l2(P.ep + pe);
            if ((lane & 7) == 2) prefetch_l2(P.tinfo + pe);
            if (!MANY && (lane & 3) == 3) prefetch_l2(reinterpret_cast<const char*>(actions) + (int64_t)pe * P.action_bytes);
        }
    }
#endif

    const uint64_t* zig = use_zig ? zig_kiwi : nullptr;
    if (!MANY) {
        if (FLAT) step_one_flat<MOD>(P, io, action_raw, adtype, true, e, active, lane, warp_env0, nvalid, stage[wid], zig,
                                     plan_lut, io.obs, io.reward, io.done, i, j, k, meta, tinfo, ep, ep_ret);
        else step_one<NV, MOD, EVAL, PAC>(P, io, action_raw, adtype, true, e, active, lane, warp_env0, nvalid, stage[wid],
                                          zig, plan_lut, io.obs, io.reward, io.done, i, j, k, meta, tinfo, ep, ep_ret);
    } else {
        for (int t = 0; t < T; ++t) {
            const long long a_now = action_raw;
            if (t + 1 < T) action_raw = load_action_raw(actions, adtype, (int64_t)(t + 1) * n_envs + le);   // next step's
            float* obs_t = io.obs + (int64_t)t * P.obs_elems;
            float* rew_t = io.reward + (int64_t)t * n_envs;
            uint8_t* done_t = io.done + (int64_t)t * n_envs;
            if (FLAT) step_one_flat<MOD>(P, io, a_now, adtype, false, e, active, lane, warp_env0, nvalid, stage[wid], zig,
                                         plan_lut, obs_t, rew_t, done_t, i, j, k, meta, tinfo, ep, ep_ret);
            else step_one<NV, MOD, false, PAC>(P, io, a_now, adtype, false, e, active, lane, warp_env0, nvalid, stage[wid],
                                               zig, plan_lut, obs_t, rew_t, done_t, i, j, k, meta, tinfo, ep, ep_ret);
        }
    }
    if (active) {
        PTG_ST_STATE(P.core + e, make_int4(i, j, k, (int)meta));
        PTG_ST_STATE(P.tinfo + e, tinfo);
        PTG_ST_STATE(P.ep_ret + e, ep_ret);
    }
    }
    if (elect_one()) tma_store_wait_read();             // the staging buffer must outlive the bulk stores' reads
}

// ------------------------------------------------------------------------------------------------------------
// episode statistics: deterministic reduction in one launch
// ------------------------------------------------------------------------------------------------------------
struct StatAcc { double cnt, sum, sq, len, mn, mx; };

__device__ __forceinline__ StatAcc stat_combine(const StatAcc& a, const StatAcc& b) {
    return StatAcc{a.cnt + b.cnt, a.sum + b.sum, a.sq + b.sq, a.len + b.len, fmin(a.mn, b.mn), fmax(a.mx, b.mx)};
}
__device__ __forceinline__ StatAcc stat_shfl_down(const StatAcc& a, int o) {
    return StatAcc{__shfl_down_sync(0xffffffffu, a.cnt, o), __shfl_down_sync(0xffffffffu, a.sum, o),
                   __shfl_down_sync(0xffffffffu, a.sq, o), __shfl_down_sync(0xffffffffu, a.len, o),
                   __shfl_down_sync(0xffffffffu, a.mn, o), __shfl_down_sync(0xffffffffu, a.mx, o)};
}
__device__ __forceinline__ StatAcc stat_block_reduce(StatAcc a) {
    __shared__ StatAcc sm[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a = stat_combine(a, stat_shfl_down(a, o));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sm[wid] = a;
    __syncthreads();
    const int nw = blockDim.x >> 5;
    a = (threadIdx.x < nw) ? sm[threadIdx.x] : StatAcc{0, 0, 0, 0, INFINITY, -INFINITY};
    if (wid == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a = stat_combine(a, stat_shfl_down(a, o));
    }
    return a;   // valid in thread 0
}

#define PTG_STATS_BLOCKS 592   // 4 x 148 SMs

// One launch: every CTA reduces its blocked env range to a partial; the LAST CTA to finish (atomic ticket) folds the
// partials in index order -- the same fixed summation order whichever CTA that is -- and writes the record.
__global__ void k_episode_stats(const __grid_constant__ DevParams P, StatAcc* partial, unsigned int* ticket, int clear,
                                double total_steps, PtgEpisodeStats* out) {
    StatAcc a{0, 0, 0, 0, INFINITY, -INFINITY};
    // fixed env -> thread assignment (blocked ranges) so that the summation order never depends on timing
    const int64_t per_block = (P.n_envs + gridDim.x - 1) / gridDim.x;
    const int64_t lo = per_block * blockIdx.x, hi = min(P.n_envs, lo + per_block);
    for (int64_t e = lo + threadIdx.x; e < hi; e += blockDim.x) {
        const int c = P.fin_cnt[e];
        if (c > 0) {
            a = stat_combine(a, StatAcc{(double)c, P.fin_ret_sum[e], P.fin_ret_sq[e], P.fin_len_sum[e], P.fin_min[e],
                                        P.fin_max[e]});
            if (clear) {
                P.fin_cnt[e] = 0; P.fin_ret_sum[e] = 0.0; P.fin_ret_sq[e] = 0.0; P.fin_len_sum[e] = 0.0;
                P.fin_min[e] = INFINITY; P.fin_max[e] = -INFINITY;
            }
        }
    }
    a = stat_block_reduce(a);
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = a;
        __threadfence();
        is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    StatAcc t{0, 0, 0, 0, INFINITY, -INFINITY};
    for (int q = threadIdx.x; q < (int)gridDim.x; q += blockDim.x) {      // (L2 reads: the partials come from other SMs)
        const double* pq = reinterpret_cast<const double*>(partial + q);
        t = stat_combine(t, StatAcc{__ldcg(pq), __ldcg(pq + 1), __ldcg(pq + 2), __ldcg(pq + 3), __ldcg(pq + 4), __ldcg(pq + 5)});
    }
    t = stat_block_reduce(t);
    if (threadIdx.x == 0) {
        out->count = t.cnt; out->sum_return = t.sum; out->sum_return_sq = t.sq; out->sum_length = t.len;
        out->min_return = t.mn; out->max_return = t.mx; out->total_steps = total_steps; out->_reserved = 0.0;
        *ticket = 0u;
    }
}

// ------------------------------------------------------------------------------------------------------------
// state snapshot helpers (get/set state): unpack to / pack from plain SoA int32 arrays on the device
// ------------------------------------------------------------------------------------------------------------
struct StateDev {
    int32_t *meth_state, *i, *j, *k, *hot_cold, *standby_ds, *startup_ds, *partial_ds, *full_ds, *current_action,
            *act_ep_h, *act_ep_d, *episode_count;
    int64_t* draws;
    double *t_cat, *cum_reward;
    uint64_t* rng;            // [n][4] = {state_hi, state_lo, inc_hi, inc_lo}
    uint32_t* state_changes;
};

__global__ void k_state_unpack(const __grid_constant__ DevParams P, StateDev s, const double* vals) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= P.n_envs) return;
    const int4 c = P.core[e];
    const Meta m = meta_unpack((uint32_t)c.w);
    s.meth_state[e] = m.state; s.i[e] = c.x; s.j[e] = c.y; s.k[e] = c.z; s.hot_cold[e] = m.hot_cold;
    s.standby_ds[e] = m.standby_ds;
    s.startup_ds[e] = m.startup_ds;
    s.partial_ds[e] = m.part_ds; s.full_ds[e] = m.full_ds; s.current_action[e] = m.cur_action;
    const int2 ep = P.ep[e];
    s.act_ep_h[e] = ep.x; s.act_ep_d[e] = ep.y; s.episode_count[e] = P.ep_count[e];
    s.draws[e] = P.draws_total[e] + (int64_t)ti_draws(P.tinfo[e]);
    const RngRec rr = P.rng[e];
    s.rng[4 * e] = rr.s_hi; s.rng[4 * e + 1] = rr.s_lo; s.rng[4 * e + 2] = rr.i_hi; s.rng[4 * e + 3] = rr.i_lo;
    s.state_changes[e] = P.has_penalty ? P.nchg[e] : 0u;
    s.t_cat[e] = vals[ti_id(P.tinfo[e])];
    s.cum_reward[e] = P.ep_ret[e];
}

__global__ void k_state_pack(const __grid_constant__ DevParams P, StateDev s, const BuildParams B) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= P.n_envs) return;
    Meta m;
    m.state = s.meth_state[e]; m.hot_cold = s.hot_cold[e];
    m.standby_ds = s.standby_ds[e]; m.startup_ds = s.startup_ds[e];
    m.part_ds = s.partial_ds[e]; m.full_ds = s.full_ds[e]; m.cur_action = s.current_action[e];
    P.core[e] = make_int4(s.i[e], s.j[e], s.k[e], (int)meta_pack(m));
    P.ep[e] = make_int2(s.act_ep_h[e], s.act_ep_d[e]);
    P.ep_count[e] = s.episode_count[e];
    P.draws_total[e] = s.draws[e];                // (the 12-bit counter in tinfo restarts at 0 below)
    RngRec rr;
    rr.s_hi = s.rng[4 * e]; rr.s_lo = s.rng[4 * e + 1]; rr.i_hi = s.rng[4 * e + 2]; rr.i_lo = s.rng[4 * e + 3];
    P.rng[e] = rr;
    if (P.has_penalty) P.nchg[e] = s.state_changes[e];
    P.tinfo[e] = tinfo_of(B, s.t_cat[e]);        // t_cat must be a temperature present in the tables (or 16)
    P.ep_ret[e] = s.cum_reward[e];
}

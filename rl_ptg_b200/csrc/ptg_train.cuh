// ptg_train.cuh -- the callers either side of the env path, on the device:
//
//   k_vecnorm_*    SB3 VecNormalize(norm_obs=False) reward normalisation (src/rl_utils.py:453): discounted-return
//                  update, batch moments (deterministic Welford/Chan tree: warp shuffles -> per-CTA -> one CTA),
//                  RunningMeanStd.update_from_moments, clip(reward / sqrt(var + eps))
//   k_features     the flat feature rows SB3's CombinedExtractor builds from the Dict observation (keys in sorted
//                  order, METH_STATUS one-hot(6)): [n_envs, 40] (mod) / [n_envs, 31] (raw) fp32, one bulk store/warp
//   k_gae          RolloutBuffer.compute_returns_and_advantage (GAE(lambda)) as a backward scan per env
//
// All three are HBM-bound streaming kernels; one env per thread, coalesced accesses.
#pragma once
#include "ptg_kernels.cuh"

// ------------------------------------------------------------------------------------------------------------
// VecNormalize (reward)
// ------------------------------------------------------------------------------------------------------------
struct Moments { double n, mean, m2; };     // count, mean, sum of squared deviations

// Chan et al. pairwise combine (the same algebra as RunningMeanStd.update_from_moments)
__device__ __forceinline__ Moments moments_combine(const Moments& a, const Moments& b) {
    if (b.n == 0.0) return a;
    if (a.n == 0.0) return b;
    const double n = a.n + b.n, d = b.mean - a.mean;
    return Moments{n, a.mean + d * b.n / n, a.m2 + b.m2 + d * d * a.n * b.n / n};
}
__device__ __forceinline__ Moments moments_shfl_down(const Moments& a, int o) {
    return Moments{__shfl_down_sync(0xffffffffu, a.n, o), __shfl_down_sync(0xffffffffu, a.mean, o),
                   __shfl_down_sync(0xffffffffu, a.m2, o)};
}
__device__ __forceinline__ Moments moments_block_reduce(Moments a) {
    __shared__ Moments sm[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a = moments_combine(a, moments_shfl_down(a, o));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sm[wid] = a;
    __syncthreads();
    const int nw = blockDim.x >> 5;
    a = (threadIdx.x < nw) ? sm[threadIdx.x] : Moments{0.0, 0.0, 0.0};
    if (wid == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a = moments_combine(a, moments_shfl_down(a, o));
    }
    return a;   // valid in thread 0
}

// returns = returns * gamma + reward (VecNormalize._update_reward) and the moments of the new returns.
// Every thread accumulates {count, sum, sum of squares} of (x - pivot) with the SAME pivot (the running mean so far:
// the returns scatter around it, so there is no cancellation), which makes the partial results plainly additive: the
// warp / CTA / grid trees are fixed-order additions (bit-reproducible, no divisions), and one thread converts the
// grand total to {n, mean, M2} at the end.  The LAST CTA to finish (atomic ticket) folds the per-CTA partials, so
// the whole reduction is one launch.
struct Sums { double n, s1, s2; };
__device__ __forceinline__ Sums sums_add(const Sums& a, const Sums& b) { return Sums{a.n + b.n, a.s1 + b.s1, a.s2 + b.s2}; }
__device__ __forceinline__ Sums sums_block_reduce(Sums a) {
    __shared__ Sums sm[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        a = sums_add(a, Sums{__shfl_down_sync(0xffffffffu, a.n, o), __shfl_down_sync(0xffffffffu, a.s1, o),
                             __shfl_down_sync(0xffffffffu, a.s2, o)});
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sm[wid] = a;
    __syncthreads();
    const int nw = blockDim.x >> 5;
    a = (threadIdx.x < nw) ? sm[threadIdx.x] : Sums{0.0, 0.0, 0.0};
    if (wid == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            a = sums_add(a, Sums{__shfl_down_sync(0xffffffffu, a.n, o), __shfl_down_sync(0xffffffffu, a.s1, o),
                                 __shfl_down_sync(0xffffffffu, a.s2, o)});
    }
    __syncthreads();
    return a;   // valid in thread 0
}

__global__ void __launch_bounds__(256) k_vecnorm_returns(int64_t n, const float* __restrict__ reward,
                                                         double* __restrict__ returns, double gamma,
                                                         const double* __restrict__ st_in, Sums* __restrict__ partial,
                                                         unsigned int* __restrict__ ticket, Moments* __restrict__ out) {
    const double pivot = st_in[0];
    Sums a{0.0, 0.0, 0.0};
    const int64_t per_block = (((n + gridDim.x - 1) / gridDim.x) + 1) & ~(int64_t)1;      // even: 16 B aligned pairs
    const int64_t lo = per_block * blockIdx.x, hi = min(n, lo + per_block);
    for (int64_t e = lo + 2 * threadIdx.x; e < hi; e += 2 * blockDim.x) {
        if (e + 1 < hi) {
            double2 r = *reinterpret_cast<const double2*>(returns + e);
            const float2 w = *reinterpret_cast<const float2*>(reward + e);
            r.x = r.x * gamma + (double)w.x;
            r.y = r.y * gamma + (double)w.y;
            *reinterpret_cast<double2*>(returns + e) = r;
            const double dx = r.x - pivot, dy = r.y - pivot;
            a.n += 2.0; a.s1 += dx; a.s1 += dy; a.s2 += dx * dx; a.s2 += dy * dy;
        } else {
            const double r = returns[e] * gamma + (double)reward[e];
            returns[e] = r;
            const double dx = r - pivot;
            a.n += 1.0; a.s1 += dx; a.s2 += dx * dx;
        }
    }
    a = sums_block_reduce(a);
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = a;
        __threadfence();
        is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    Sums b{0.0, 0.0, 0.0};            // sequential per thread over a strided subset, then the fixed tree
    for (int q = threadIdx.x; q < (int)gridDim.x; q += blockDim.x) {
        const double* pq = reinterpret_cast<const double*>(partial + q);      // other CTAs' results: read at L2
        b = sums_add(b, Sums{__ldcg(pq), __ldcg(pq + 1), __ldcg(pq + 2)});
    }
    b = sums_block_reduce(b);
    if (threadIdx.x == 0) {
        const double m = b.n > 0.0 ? b.s1 / b.n : 0.0;
        *out = Moments{b.n, pivot + m, b.s2 - b.s1 * m};
        *ticket = 0u;
    }
}

// RunningMeanStd.update_from_moments with the batch moments of every rank (rank order), then
// normalize_reward: clip(reward / sqrt(var + epsilon), -clip, clip); returns[done] = 0.
// st_in / st_out are distinct buffers ({mean, var, count, pad}): every CTA reads the old statistics, CTA 0 writes the
// new ones.
__global__ void __launch_bounds__(256) k_vecnorm_apply(int64_t n, const float* __restrict__ reward_in,
                                                       const uint8_t* __restrict__ done, double* __restrict__ returns,
                                                       const double* __restrict__ st_in, double* __restrict__ st_out,
                                                       const Moments* __restrict__ batch, int n_batch, int training,
                                                       double epsilon, double clip, float* __restrict__ reward_out) {
    // the statistics update is a few dozen dependent fp64 operations: once per CTA, not once per env
    __shared__ double s_std;
    if (threadIdx.x == 0) {
        double mean = st_in[0], var = st_in[1], count = st_in[2];
        if (training) {
            Moments b{0.0, 0.0, 0.0};
            for (int r = 0; r < n_batch; ++r) b = moments_combine(b, batch[r]);
            const double batch_var = b.m2 / b.n;                       // np.var (population)
            const double delta = b.mean - mean, tot = count + b.n;
            const double new_mean = mean + delta * b.n / tot;
            const double m_2 = var * count + batch_var * b.n + delta * delta * count * b.n / tot;
            mean = new_mean; var = m_2 / tot; count = tot;
        }
        if (blockIdx.x == 0) { st_out[0] = mean; st_out[1] = var; st_out[2] = count; st_out[3] = 0.0; }
        s_std = sqrt(var + epsilon);
    }
    __syncthreads();
    const double inv_std = 1.0 / s_std;      // (x * (1/s) differs from x / s by <= 1 ulp of fp64: invisible in the fp32 result)
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const double scaled = (double)reward_in[e] * inv_std;
        reward_out[e] = (float)fmin(fmax(scaled, -clip), clip);
        if (done[e]) returns[e] = 0.0;
    }
}

// ------------------------------------------------------------------------------------------------------------
// flat policy features
// ------------------------------------------------------------------------------------------------------------
// Key-major observation buffer (ptg_obs_layout) -> [n_envs, F] rows in the key order of a gymnasium Dict space /
// SB3 CombinedExtractor (sorted keys; METH_STATUS as one-hot(6)):
//   mod: CH4, Elec_Heating, H2O, H2_in, H2_res, METH_STATUS[6], Part_Full[pa], Pot_Reward[pa], T_CAT, cos, sin
//   raw: CH4, EUA_Price[2], Elec_Heating, Elec_Price[pa], Gas_Price[2], H2O, H2_in, H2_res, METH_STATUS[6], T_CAT,
//        cos, sin
// Each warp assembles its 32 rows in shared memory (row stride F; the scalar columns are written column-wise,
// the window columns copied from the already row-major window blocks) and stores them with one bulk copy.
#define PTG_FEAT_MAX 48
__global__ void __launch_bounds__(PTG_BLOCK) k_features(const __grid_constant__ DevParams P, const float* __restrict__ obs,
                                                        float* __restrict__ feat, int F) {
    extern __shared__ __align__(128) float fsm[];
    const int n = (int)P.n_envs;
    const int e = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int warp_env0 = e - lane;
    if (warp_env0 >= n) return;
    const int nvalid = min(32, n - warp_env0);
    float* tile = fsm + wid * 32 * F;
    const int pa = P.pa;
    const bool mod = !P.raw;
    // column map (row-relative positions)
    const int c_ch4 = 0, c_heat = mod ? 1 : 3, c_h2o = mod ? 2 : 4 + pa + 2, c_h2in = c_h2o + 1, c_h2res = c_h2o + 2;
    const int c_oh = c_h2res + 1;
    const int c_win0 = mod ? c_oh + 6 + pa : 4;           // Pot_Reward (mod) | Elec_Price (raw)
    const int c_win1 = c_oh + 6;                          // Part_Full (mod only)
    const int c_T = mod ? c_oh + 6 + 2 * pa : c_oh + 6;
    if (lane < nvalid) {
        const float* sc = obs + P.off_scalar + e;
        float* row = tile + lane * F;
        const int status = __float_as_int(sc[0]);
        // scalar block order of the obs buffer: status, T, H2, CH4, H2_res, H2O, heat, sin, cos
        row[c_T] = sc[1 * P.n_pad];
        row[c_h2in] = sc[2 * P.n_pad];
        row[c_ch4] = sc[3 * P.n_pad];
        row[c_h2res] = sc[4 * P.n_pad];
        row[c_h2o] = sc[5 * P.n_pad];
        row[c_heat] = sc[6 * P.n_pad];
        row[c_T + 2] = sc[7 * P.n_pad];       // sin
        row[c_T + 1] = sc[8 * P.n_pad];       // cos
#pragma unroll
        for (int s = 0; s < 6; ++s) row[c_oh + s] = (status == s) ? 1.0f : 0.0f;
        if (!mod) {
            const float2 gas = reinterpret_cast<const float2*>(obs + P.off_gas)[e];
            const float2 eua = reinterpret_cast<const float2*>(obs + P.off_eua)[e];
            row[1] = eua.x; row[2] = eua.y;
            row[4 + pa] = gas.x; row[4 + pa + 1] = gas.y;
        }
    }
    // window blocks: [n_envs][pa] row-major in the obs buffer -> coalesced reads of the warp's 32*pa values
    const float* w0 = obs + P.off_win0 + (int64_t)warp_env0 * pa;
    const float* w1 = obs + P.off_win1 + (int64_t)warp_env0 * pa;
    for (int idx = lane; idx < nvalid * pa; idx += 32) {
        const int r = idx / pa, a = idx - r * pa;
        tile[r * F + c_win0 + a] = w0[idx];
        if (mod) tile[r * F + c_win1 + a] = w1[idx];
    }
    float* g = feat + (int64_t)warp_env0 * F;
    const uint32_t bytes = (uint32_t)(nvalid * F * 4);
    if ((bytes & 15u) == 0 && (((uintptr_t)g) & 15u) == 0) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            tma_store_1d(g, tile, bytes);
            tma_store_commit();
            tma_store_wait_read();
        }
    } else {                                  // ragged tail / odd alignment: plain stores
        __syncwarp();
        for (int idx = lane; idx < nvalid * F; idx += 32) g[idx] = tile[idx];
    }
}

// ------------------------------------------------------------------------------------------------------------
// GAE(lambda): SB3 RolloutBuffer.compute_returns_and_advantage in numpy's evaluation order and dtypes
// ------------------------------------------------------------------------------------------------------------
//   next_non_terminal = 1 - (step == T-1 ? dones : episode_starts[step+1])
//   delta = rewards[step] + gamma * next_values * next_non_terminal - values[step]
//   last  = delta + gamma * lambda * next_non_terminal * last
//   advantages[step] = last ; returns = advantages + values
// numpy's promotion rules make this a mixed-precision scan: the buffers are float32, `dones` is a bool array so
// `1.0 - dones` (last step only) is float64, and `last_gae_lam` is float64 from then on; `delta` of the earlier steps
// is pure float32.  The kernel follows that literally (one env per thread, backward over T, coalesced across envs).
__global__ void k_gae(int64_t n, int T, const float* __restrict__ rewards, const float* __restrict__ values,
                      const uint8_t* __restrict__ episode_starts, const float* __restrict__ last_values,
                      const uint8_t* __restrict__ last_dones, double gamma, double gae_lambda,
                      float* __restrict__ advantages, float* __restrict__ returns) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const float gamma_f = (float)gamma, gl_f = (float)(gamma * gae_lambda);
    double last = 0.0;
    {   // step T-1: float64 next_non_terminal
        const int64_t q = (int64_t)(T - 1) * n + e;
        const double nnt = 1.0 - (double)last_dones[e];
        const float v = values[q];
        const double delta = ((double)rewards[q] + (double)__fmul_rn(gamma_f, last_values[e]) * nnt) - (double)v;
        last = delta + ((gamma * gae_lambda) * nnt) * last;
        const float adv = (float)last;
        advantages[q] = adv;
        returns[q] = __fadd_rn(adv, v);
    }
#pragma unroll 8
    for (int t = T - 2; t >= 0; --t) {      // (unrolled: the loads of 8 steps are in flight together)
        const int64_t q = (int64_t)t * n + e, qn = q + n;
        const float nnt = __fsub_rn(1.0f, (float)episode_starts[qn]);
        const float v = values[q];
        const float delta = __fsub_rn(__fadd_rn(rewards[q], __fmul_rn(__fmul_rn(gamma_f, values[qn]), nnt)), v);
        last = (double)delta + (double)__fmul_rn(gl_f, nnt) * last;
        const float adv = (float)last;
        advantages[q] = adv;
        returns[q] = __fadd_rn(adv, v);
    }
}

// ------------------------------------------------------------------------------------------------------------
// calculate_optimum (src/rl_opt.py:26-152): the per-hour theoretical optimum T-OPT whose columns 20 / 23 become the
// Pot_Reward / Part_Full observations
// ------------------------------------------------------------------------------------------------------------
// Everything that depends only on the load level (partial / full) -- flows, the PEM efficiency polynomial, the
// price-independent revenues -- is evaluated by the host in the reference's operation order (two levels); the kernel
// does the per-hour part for both levels, picks the better one (ties -> partial load, like list.index(max(...))) and
// writes the reference's (n_hours, 24) statistics row.  Column 21 (the running sum) is accumulated strictly
// sequentially like the reference's `cum_rew += rew`, by one thread of a second launch.
struct OptParams {
    PtgOptLevel lv[3];      // off, partial load, full load
};

__global__ void k_calculate_optimum(const double* __restrict__ el, int64_t n_hours, const double* __restrict__ gas,
                                    const double* __restrict__ eua, int64_t n_days,
                                    const __grid_constant__ OptParams O, double* __restrict__ stats) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_hours) return;
    int64_t t_day = t / 24;
    if (t_day == n_days) t_day -= 1;                                   // rl_opt.py:52
    const double e = el[t], g = gas[t_day], u = eua[t_day];
    double rew_l[2], part[2][8];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        const PtgOptLevel& L = O.lv[l + 1];
        const double ch4_rev = L.q_gas * g;
        const double eua_rev = L.k_eua * u * 100;
        const double heat_cost = L.k_heat * e;
        const double ely_cost = L.k_ely * e;
        const double elec_costs = heat_cost + ely_cost;
        rew_l[l] = ch4_rev + L.chp_rev + L.steam_rev + eua_rev + L.o2_rev - elec_costs - L.water_cost;
        part[l][0] = ch4_rev; part[l][1] = L.steam_rev; part[l][2] = L.o2_rev; part[l][3] = eua_rev;
        part[l][4] = L.chp_rev; part[l][5] = -heat_cost; part[l][6] = -ely_cost; part[l][7] = -L.water_cost;
    }
    const int index = rew_l[1] > rew_l[0] ? 1 : 0;
    const double rew = rew_l[index];
    const bool on = rew > 0;
    double* r = stats + t * 24;
    r[0] = (double)t; r[1] = e; r[2] = g; r[3] = u;
    const PtgOptLevel& S = O.lv[on ? index + 1 : 0];
#pragma unroll
    for (int q = 0; q < 8; ++q) r[4 + q] = S.stat8[q];
#pragma unroll
    for (int q = 0; q < 8; ++q) r[12 + q] = on ? part[1][q] : 0.0;     // reference quirk: always the full-load values
    r[20] = rew;
    r[21] = on ? rew : 0.0;                                            // summed up by k_optimum_cumsum
    r[22] = 0.0;
    r[23] = on ? (double)index : -1.0;
}

__global__ void k_optimum_cumsum(int64_t n_hours, double* __restrict__ stats) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double cum = 0.0;
    for (int64_t t = 0; t < n_hours; ++t) { cum += stats[t * 24 + 21]; stats[t * 24 + 21] = cum; }
}

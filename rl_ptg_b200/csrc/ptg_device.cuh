// ptg_device.cuh -- device-side data layout and the per-env step of the batched PtG environment.
//
// One env per thread.  Everything PTGEnv.step() recomputes from the experimental tables (env/ptg_gym_env.py:
// 452-458: T_cat = op[-1,1], five np.average over the S-row window) is a pure function of (table, start row)
// and lives in a pre-built *step table*; everything _get_index computes (:514-523) is a pure function of
// (target table, T_cat) where T_cat only ever takes values found in the tables and lives in the *argmin LUT*.
// Market data is re-laid as one 16B-aligned row per hour / per day.  All tables are L2-resident (< 20 MB).
#pragma once
#include <stdint.h>

#include "../../include/ptg_b200.h"
#include "ptg_rng.cuh"

#ifndef PTG_BLOCK
#define PTG_BLOCK 256            // threads per CTA of the step kernels
#endif
#define PTG_N_ARGMIN 6           // argmin targets: cooldown, standby_up, standby_down, startup_cold, startup_hot, op1

// --- step table entry: 64 B, 64 B aligned (exactly two 32 B sectors per gather) ---------------------------
// The reward of :280-334 is linear in the three prices once the window means are fixed:
//     rew = c_gas * gas + c_eua * eua - c_el * el + c_0            (already scaled by sim_step / 3600)
// so the table-build kernel evaluates the flow-dependent factors -- including the PEM electrolyzer efficiency
// basis functions (:312-317) and the CHP/EEG terms (:291-297) -- once per (table, start row) and the step
// kernel is left with three fp64 FMAs.
struct __align__(16) StepEntry {
    double c_gas, c_eua, c_el, c_0;
    float norm[6];      // min-max normalised T_cat, H2, CH4, H2_res, H2O, P_el -- fp32(fp64 expression of :212-217)
    int32_t tinfo;      // (T value id << 3) | flags(T_end)
    int32_t _pad;
};
static_assert(sizeof(StepEntry) == 64, "StepEntry layout");

// --- window statistics in fp64, only read for the eval-mode info dict (:251-278) -------------------------
struct __align__(16) StepMeans {
    double mean[5];     // window means: H2, CH4, H2_res, H2O, P_el (fp64, numpy pairwise order)
    double t_end;       // T_cat after the step (last row of the window)
};

#define PTG_TF_COLD 1       // T <= t_cat_startup_cold   (:339)
#define PTG_TF_HOT 2        // T >= t_cat_startup_hot    (:341)
#define PTG_TF_SBUP 4       // T <= t_cat_standby        (:579)
// per-env `tinfo` word: [0,3) flags | [3,20) T value id | [20,32) noise draws since the last fold into RngRec
#define PTG_TI_ID_BITS 17
#define PTG_TI_DRAW_SHIFT 20
#define PTG_TI_DRAW_MAX 0xfffu
#define PTG_TI_LOW_MASK 0xfffffu
PTG_HD int ti_id(int32_t t) { return (int)(((uint32_t)t >> 3) & ((1u << PTG_TI_ID_BITS) - 1u)); }
PTG_HD uint32_t ti_draws(int32_t t) { return (uint32_t)t >> PTG_TI_DRAW_SHIFT; }

// --- day row: 32 B ------------------------------------------------------------------------------------------
struct __align__(16) DayRow {
    double gas, eua;                       // g_e[0,0,d], g_e[1,0,d]  (reward, info)
    float gas_n0, gas_n1, eua_n0, eua_n1;  // normalised (today, tomorrow), raw observation design
};

// --- per-step clock table, indexed by k+1 ---------------------------------------------------------------------
struct __align__(16) ClockRow {
    float sin_h, cos_h;     // fp32(math.sin/cos(2*pi*clock_hours)), :449-450
    int32_t h_step, d_step; // floor(clock_hours), floor(clock_days), :444-445
};

// --- packed per-env plant state ---------------------------------------------------------------------------------
// core = {i, j, k, meta};  meta bits: [0,3) Meth_State | 3 hot_cold | [4,7) current_action | [7,32) "tabs": five
// 5-bit table ids indexed by state id = the table currently bound to self.standby / cooldown / self.startup /
// self.partial / self.full (env/ptg_gym_env.py:111-116), so "the table of the current state" is one shift.
#define PTG_META_TABS_SHIFT 7
PTG_HD uint32_t meta_tab(uint32_t meta, int state) { return (meta >> (PTG_META_TABS_SHIFT + 5 * state)) & 31u; }
PTG_HD uint32_t meta_set_tab(uint32_t meta, int state, uint32_t ds) {
    const int sh = PTG_META_TABS_SHIFT + 5 * state;
    return (meta & ~(31u << sh)) | (ds << sh);
}
struct Meta {     // unpacked view, cold paths only
    int state, hot_cold, cur_action, standby_ds, startup_ds, part_ds, full_ds;
};
PTG_HD Meta meta_unpack(uint32_t m) {
    Meta r;
    r.state = m & 7; r.hot_cold = (m >> 3) & 1; r.cur_action = (m >> 4) & 7;
    r.standby_ds = meta_tab(m, PTG_STANDBY); r.startup_ds = meta_tab(m, PTG_STARTUP);
    r.part_ds = meta_tab(m, PTG_PARTIAL_LOAD); r.full_ds = meta_tab(m, PTG_FULL_LOAD);
    return r;
}
PTG_HD uint32_t meta_pack(const Meta& r) {
    uint32_t m = (uint32_t)r.state | ((uint32_t)r.hot_cold << 3) | ((uint32_t)r.cur_action << 4);
    m = meta_set_tab(m, PTG_STANDBY, r.standby_ds);
    m = meta_set_tab(m, PTG_COOLDOWN, PTG_DS_COOLDOWN);
    m = meta_set_tab(m, PTG_STARTUP, r.startup_ds);
    m = meta_set_tab(m, PTG_PARTIAL_LOAD, r.part_ds);
    return meta_set_tab(m, PTG_FULL_LOAD, r.full_ds);
}

// load-change chains of _partial / _full (:636-688, :702-754) as look-up tables over time_op, built on the host
// from the config thresholds: entry = table id | i/j rule | new i
#define PTG_CHAIN_P_FROM_OP2F 0      // _partial after op2_start_f
#define PTG_CHAIN_P_FROM_OP3 1       // _partial after op3_p_f
#define PTG_CHAIN_F_FROM_OP1 2       // _full after op1_start_p
#define PTG_CHAIN_F_FROM_OP8 3       // _full after op8_f_p
#define PTG_CHAIN_J_ONE 0u           // i := new i, j := 1
#define PTG_CHAIN_J_KEEP 1u          // i kept, j += 1            (:658, :724)
#define PTG_CHAIN_J_DEVELOPED 2u     // i, j := i/j_fully_developed (:652-653, :718-719)
#define PTG_CHAIN_J_ARGMIN 3u        // i := argmin(op1_start_p, T_cat), j := 1 (:641-642)
PTG_HD uint32_t chain_pack(uint32_t ds, uint32_t rule, uint32_t new_i) { return (ds << 27) | (rule << 25) | (new_i & 0x1ffffffu); }

// Scalars of the reward model, shared by the table-build kernel and the eval-mode info path.
struct RewardConsts {
    double convert_mol_to_Nm3, H_u_CH4, H_u_H2, dt_water, cp_water, rho_water, Molar_mass_CO2, Molar_mass_H2O,
           h_H2O_evap, eeg_el_price, heat_price, o2_price, water_price, min_load_electrolyzer, max_h2_volumeflow,
           eta_CHP, sim_step_d, penalty;
    int32_t b_s3, _pad;
};

// --- per-env noise generator: ONE 32 B sector holds everything a draw needs (one 256-bit load, one 128-bit store).
//     The record is deliberately exactly one sector: 1 M envs are 32 MB, which -- accessed with the evict_last policy
//     while everything that streams carries evict_first -- stays resident in the 126 MB L2 from step to step, so a
//     draw is an L2 hit instead of a dependent DRAM round trip.  The cold 64-bit draw counter lives in its own array
//     (DevParams.draws_total) and is only touched when the 12-bit per-env counter that rides in the spare bits of
//     `tinfo` (already streamed every step) overflows ---
struct __align__(32) RngRec {
    uint64_t s_hi, s_lo;   // PCG64 state
    uint64_t i_hi, i_lo;   // PCG64 increment (constant per seed)
};
static_assert(sizeof(RngRec) == 32, "RngRec layout");

#if defined(__CUDACC__)
// --- 256-bit global accesses (sm_100+: LDG.E.256 / STG.E.256): one instruction and ONE L1 wavefront per lane for a
//     32 B sector instead of two 128-bit accesses -- the table gathers of the step kernel are wavefront-bound ---
// L2 residency: one step streams ~250 MB of per-env state and observations through the 126 MB L2, which evicts the
// (16 MB of) tables unless the table gathers ask for evict_last and the streaming stores for evict_first.  The
// policy words are what `createpolicy.fractional.L2::evict_{last,first}.b64 p, 1.0` returns (as immediates they
// cost no live registers).
#ifndef PTG_L2_HINTS
#define PTG_L2_HINTS 1
#endif
#define PTG_L2_EVICT_FIRST 0x12F0000000000000ull
#define PTG_L2_EVICT_LAST 0x14F0000000000000ull
struct U256 { unsigned long long a, b, c, d; };
// (256-bit accesses take the L2 eviction priority as an instruction modifier -- LDG.E.ELL2.256 -- instead of a policy
//  operand: no pair of uniform-register moves in front of every gather)
__device__ __forceinline__ U256 ldg256_nc(const void* p) {      // read-only tables (non-coherent path)
    U256 r;
#if PTG_L2_HINTS
    asm volatile("ld.global.nc.L2::evict_last.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(r.a), "=l"(r.b), "=l"(r.c), "=l"(r.d) : "l"(p));
#else
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(r.a), "=l"(r.b), "=l"(r.c), "=l"(r.d) : "l"(p));
#endif
    return r;
}
__device__ __forceinline__ int ldg32_nc_keep(const int* p) {     // small read-only tables (argmin LUT)
#if PTG_L2_HINTS
    int r;
    asm volatile("ld.global.nc.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(PTG_L2_EVICT_LAST));
    return r;
#else
    return __ldg(p);
#endif
}
// streaming stores (written once per step, re-read one full step later at the earliest)
template <typename T>
__device__ __forceinline__ void st_stream(T* p, T v) {
#if PTG_L2_HINTS
    __stcs(p, v);
#else
    *p = v;
#endif
}
__device__ __forceinline__ U256 ld256(const void* p) {          // read-write data
    U256 r;
    asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(r.a), "=l"(r.b), "=l"(r.c), "=l"(r.d) : "l"(p) : "memory");
    return r;
}
// (256-bit STORES are deliberately not used: ptxas 12.9 turned `st.global.v4.u64` into a single STG.E.64 here.)

// --- L2-resident read-write data (RNG records; per-env plant state when PTG_STATE_L2): evict_last on both sides ---
__device__ __forceinline__ U256 ld256_keep(const void* p) {
    U256 r;
#if PTG_L2_HINTS
    asm volatile("ld.global.L2::evict_last.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(r.a), "=l"(r.b), "=l"(r.c), "=l"(r.d) : "l"(p) : "memory");
#else
    r = ld256(p);
#endif
    return r;
}
__device__ __forceinline__ void st128_keep(void* p, unsigned long long a, unsigned long long b) {
#if PTG_L2_HINTS
    asm volatile("st.global.L2::cache_hint.v2.u64 [%0], {%1,%2}, %3;" ::"l"(p), "l"(a), "l"(b), "l"(PTG_L2_EVICT_LAST) : "memory");
#else
    *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(a, b);
#endif
}
__device__ __forceinline__ int4 ld_keep(const int4* p) {
    int4 r;
    asm volatile("ld.global.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(PTG_L2_EVICT_LAST) : "memory");
    return r;
}
__device__ __forceinline__ int2 ld_keep(const int2* p) {
    int2 r;
    asm volatile("ld.global.L2::cache_hint.v2.s32 {%0,%1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(PTG_L2_EVICT_LAST) : "memory");
    return r;
}
__device__ __forceinline__ int ld_keep(const int* p) {
    int r;
    asm volatile("ld.global.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(PTG_L2_EVICT_LAST) : "memory");
    return r;
}
__device__ __forceinline__ double ld_keep(const double* p) {
    double r;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(PTG_L2_EVICT_LAST) : "memory");
    return r;
}
__device__ __forceinline__ void st_keep(int4* p, int4 v) {
    asm volatile("st.global.L2::cache_hint.v4.s32 [%0], {%1,%2,%3,%4}, %5;"
                 ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(PTG_L2_EVICT_LAST) : "memory");
}
__device__ __forceinline__ void st_keep(int* p, int v) {
    asm volatile("st.global.L2::cache_hint.s32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(PTG_L2_EVICT_LAST) : "memory");
}
__device__ __forceinline__ void st_keep(double* p, double v) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(PTG_L2_EVICT_LAST) : "memory");
}
#endif

// --- everything a kernel needs, passed by value (__grid_constant__) ---------------------------------------------
struct DevParams {
    // sizes
    int64_t n_envs, env_id_offset, n_envs_global;
    int32_t S;                 // rows per env step = int(sim_step / time_step_op), :66
    int32_t pa;                // price_ahead
    int32_t nv;                // 16-byte vectors per hour row
    int32_t obs_dim;
    int32_t eps_sim_steps, sim_step;
    int32_t n_hours, n_days, n_vals, n_eps_ind;
    int32_t n_eps_loops;       // SUBPROC schedule: start offsets are drawn from [0, n_eps_loops), :43-44
    int32_t raw;               // 1 = raw observation design
    int32_t flat;              // 1 = flat feature-row observation layout (PTG_OBS_FLAT)
    int32_t continuous, eval_mode, noise_mode, schedule_mode;
    int32_t has_penalty;
    int32_t auto_reset;            // 1: SB3 VecEnv auto-reset inside step | 0: the env stays terminal until ptg_reset
    int32_t prefetch_distance;     // envs between a CTA and the CTA whose state it prefetches into L2
    int32_t action_bytes;          // element size of the action tensor of the current launch
    // tables (device)
    const StepEntry* step_tab;
    const StepMeans* mean_tab;     // same indexing as step_tab
    const int32_t* argmin_lut;     // [n_vals][6]
    const float4* hour_tab;        // [n_hours][nv]
    const DayRow* day_tab;         // [n_days]
    const float* hourq_tab;        // [n_hours][32] quad hour rows incl. the day's prices (k_build_hourq_tab) or nullptr
    const ClockRow* clock_tab;     // [eps_sim_steps + 1]
    const int64_t* eps_ind;        // [n_eps_ind] or nullptr
    const uint16_t* plan_lut;      // [PTG_PLAN_LUT_SIZE] plan_pack() of every (action, state, T flags, hot_cold)
    const uint32_t* chain_tab;     // [4][chain_top + 1] chain_pack() entries
    int32_t chain_top;             // time_op values >= chain_top share the last entry
    const double* pot0;            // e_r_b[1, 0, :] fp64 (info "Pot_Reward")
    const double* pf0;             // e_r_b[2, 0, :] fp64 (info "Part_Full")
    ZigTables zig;
    int32_t ent_off[PTG_N_DATASETS];   // first entry of table d in step_tab (table d has len+1 entries)
    int32_t ds_len[PTG_N_DATASETS];
    // reset constants (:105-138)
    int32_t reset_i, reset_tinfo;
    float reset_norm[6];
    double reset_flow[5];
    // thresholds (the load-change chains live in chain_tab)
    int32_t i_fully_developed, j_fully_developed;
    // scalars
    double noise, eps_len_d, penalty /* r_0 * state_change_penalty */;
    RewardConsts rc;
    double prob_thre[6];           // continuous-action thresholds, :151-155
    // obs block offsets (elements) inside the obs buffer, see ptg_obs_layout
    int64_t off_win0, off_win1;    // mod: Pot_Reward, Part_Full | raw: Elec_Price, (unused)
    int64_t off_gas, off_eua;      // raw only
    int64_t off_scalar;            // METH_STATUS block; the 8 fp32 scalar blocks follow, each padded to n_pad
    int64_t n_pad;                 // n_envs rounded up to a multiple of 4 (16 B alignment of every block)
    int64_t obs_elems;             // total fp32 elements of one obs buffer
    // per-env state (device, SoA)
    int4* core;                    // i, j, k, meta
    int32_t* tinfo;                // flags | T value id << 3 | 12-bit draw counter << 20 (ti_id / ti_draws)
    int2* ep;                      // act_ep_h, act_ep_d
    double* ep_ret;                // Monitor-style running return (sum of returned rewards)
    int32_t* ep_count;             // constructor/resets consumed (m)
    int32_t* ep_start;             // SUBPROC schedule: start offset
    uint32_t* nchg;                // state changes this episode (only maintained when has_penalty)
    RngRec* rng;                   // per-env generator record (one 32 B sector, L2-resident)
    int64_t* draws_total;          // draws folded in from the tinfo counter so far (tape position = this + counter); cold
    uint32_t step_serial;          // serial number of the current ptg_step launch (PtgIO.windows_changed stamp)
    const double* tape;            // tape-mode noise [n_envs][tape_len]
    int64_t tape_len;
    // finished-episode accumulators
    int32_t* fin_cnt;
    double* fin_ret_sum; double* fin_ret_sq; double* fin_len_sum; double* fin_min; double* fin_max;
    uint32_t* err;                 // sticky device error bits
};

#define PTG_EBIT_ACTION 1u
#define PTG_EBIT_RANGE 2u
#define PTG_EBIT_TAPE 4u
#define PTG_EBIT_PARTFULL 8u
#define PTG_EBIT_BOUNDS 16u     // PTG_DEBUG_BOUNDS builds: a table index left its table (site id in the high half)

// -DPTG_DEBUG_BOUNDS: every table index of the hot path is checked against its table; a violation sets the sticky
// error word (PTG_EBIT_BOUNDS | site << 16), clamps the index and is reported by ptg_poll_error.  The stand-in for
// compute-sanitizer (closed on the GPU pool): the GPU suite runs once against this build (tools/debug_bounds.sh).
#ifdef PTG_DEBUG_BOUNDS
#define PTG_CHECK_INDEX(P, idx, n, site)                                                        \
    do {                                                                                        \
        if ((long long)(idx) < 0 || (long long)(idx) >= (long long)(n)) {                       \
            atomicOr((P).err, PTG_EBIT_BOUNDS | ((uint32_t)(site) << 16));                      \
            (idx) = (idx) < 0 ? 0 : (n) - 1;                                                    \
        }                                                                                       \
    } while (0)
#else
#define PTG_CHECK_INDEX(P, idx, n, site) do { } while (0)
#endif

// What one step produces besides the new state (all in registers).
struct StepOut {
    float reward;
    int done;
    int status;         // METH_STATUS
    float norm[6];      // T_CAT, H2, CH4, H2_res, H2O, heating
    float sin_h, cos_h;
    int32_t t_hour, t_day;
};

// Reward constituents (eval-mode info)
struct RewardParts {
    double ch4_rev, steam_rev, o2_rev, eua_rev, chp_rev, heat_cost, ely_cost, water_cost, rew, rew_unpenalised;
};

#if defined(__CUDACC__)

// The RNG record of a drawing lane: state and increment share one 32 B sector -> ONE 256-bit request per draw.
struct RngLoad {
    U256 lo;            // {s_hi, s_lo, i_hi, i_lo}
};
__device__ __forceinline__ RngLoad request_rng(const DevParams& P, int64_t e) {
    RngLoad r;
    r.lo = ld256_keep(P.rng + e);
    return r;
}

// One draw of np_random.normal(0, noise, size=1)[0].  (Measured: inlining beats an out-of-line call here -- the
// call's register save/restore costs more than the 128-bit PCG64 arithmetic adds to the live set.)
// draws_ep: the env's 12-bit draw counter (from tinfo), incremented here and folded into the record when it is full.
__device__ __forceinline__ double draw_noise(const DevParams& P, int64_t e, const uint64_t* zig_kiwi, const RngLoad& r,
                                             uint32_t& draws_ep) {
    RngRec* rec = P.rng + e;
    double out;
    if (P.noise_mode == PTG_NOISE_TAPE) {
        const int64_t d = P.draws_total[e] + (int64_t)draws_ep;
        if (d >= P.tape_len) { atomicOr(P.err, PTG_EBIT_TAPE); out = 0.0; }
        else out = P.tape[e * P.tape_len + d];
    } else {
        Pcg64 g = {r.lo.a, r.lo.b, r.lo.c, r.lo.d};
        ZigTables zt = P.zig;
        zt.kiwi = zig_kiwi;         // the CTA's shared-memory copy of the hot {ki, wi} pairs
        const double z = pcg64_standard_normal(g, zt);
        st128_keep(rec, g.s_hi, g.s_lo);
        out = 0.0 + P.noise * z;    // random_normal: loc + scale * standard_normal
    }
    if (++draws_ep == PTG_TI_DRAW_MAX) {            // rare: fold the small counter into the 64-bit one
        P.draws_total[e] += (int64_t)PTG_TI_DRAW_MAX;
        draws_ep = 0;
    }
    return out;
}

// i = int(max(argmin + noise, 0)), :584-585
__device__ __forceinline__ int jitter_index(int idx, double nz) {
    double v = (double)idx + nz;
    if (0.0 > v) v = 0.0;
    return (int)v;
}

// _get_reward, :280-334 -- same operation order as the reference, fp64, no fused multiply-add (-fmad=false)
__device__ __forceinline__ void reward_parts(const RewardConsts& P, const double* mean, double el, double gas,
                                             double eua, int state_change, RewardParts& r) {
    const double H2 = mean[0], CH4 = mean[1], H2res = mean[2], H2O = mean[3], heat = mean[4];
    double ch4_volumeflow = CH4 * P.convert_mol_to_Nm3;
    double h2_res_volumeflow = H2res * P.convert_mol_to_Nm3;
    double Q_ch4 = ch4_volumeflow * P.H_u_CH4 * 1000;
    double Q_h2_res = h2_res_volumeflow * P.H_u_H2 * 1000;
    r.ch4_rev = (Q_ch4 + Q_h2_res) * gas;
    double power_chp = Q_ch4 * P.eta_CHP * P.b_s3;
    double Q_chp = Q_ch4 * (1 - P.eta_CHP) * P.b_s3;
    r.chp_rev = power_chp * P.eeg_el_price;
    double Q_steam = H2O * (P.dt_water * P.cp_water + P.h_H2O_evap) / 3600;
    r.steam_rev = (Q_steam + Q_chp) * P.heat_price;
    double h2_volumeflow = H2 * P.convert_mol_to_Nm3;
    double o2_volumeflow = 0.5 * h2_volumeflow * 3600;
    r.o2_rev = o2_volumeflow * P.o2_price;
    double co2 = CH4 * P.Molar_mass_CO2 / 1000;
    r.eua_rev = co2 / 1000 * 3600 * eua * 100;
    r.heat_cost = heat / 1000 * el;
    double load = h2_volumeflow / P.max_h2_volumeflow, eta;
    if (load < P.min_load_electrolyzer) {
        eta = 0.02;
    } else {   // PEM electrolyzer efficiency basis functions, :315-317
        double l2 = load * load, inv = 1.0 / load, inv2 = 1.0 / l2, inv3 = 1.0 / (l2 * load);
        eta = 0.598 - 0.325 * l2 + 0.218 * (l2 * load) + 0.01 * inv - 1.68e-3 * inv2 + 2.51e-5 * inv3;
    }
    r.ely_cost = h2_volumeflow * P.H_u_H2 * 1000 / eta * el;
    double elec_costs = r.heat_cost + r.ely_cost;
    double water_elec = H2 * P.Molar_mass_H2O / 1000 * 3600;
    r.water_cost = (H2O + water_elec) / P.rho_water * P.water_price;
    r.rew_unpenalised = (r.ch4_rev + r.chp_rev + r.steam_rev + r.eua_rev + r.o2_rev - elec_costs - r.water_cost)
                        * P.sim_step_d / 3600;
    r.rew = state_change ? r.rew_unpenalised - P.penalty : r.rew_unpenalised;
}

// The price-linear form of the reward above (see StepEntry): evaluated once per table entry by k_build_step_tab.
__device__ __forceinline__ void reward_coefficients(const RewardConsts& P, const double* mean, double& c_gas,
                                                    double& c_eua, double& c_el, double& c_0) {
    RewardParts u;
    reward_parts(P, mean, 1.0, 1.0, 1.0, 0, u);        // unit prices: each price-dependent term is its factor
    const double scale = P.sim_step_d / 3600;
    c_gas = u.ch4_rev * scale;
    c_eua = u.eua_rev * scale;
    c_el = (u.heat_cost + u.ely_cost) * scale;
    c_0 = (u.chp_rev + u.steam_rev + u.o2_rev - u.water_cost) * scale;
}

// episode schedule: which eps_ind entry does env (global id) use for its m-th constructor/reset
__device__ __forceinline__ void episode_offsets(const DevParams& P, int64_t e, int32_t m, int& ep_h, int& ep_d) {
    if (P.eps_ind == nullptr) { ep_h = 0; ep_d = 0; return; }     // val/test env, :63-64
    int64_t gid = P.env_id_offset + e, slot;
    if (P.schedule_mode == PTG_SCHED_DUMMY) slot = (P.n_envs_global * (int64_t)m + gid) % P.n_eps_ind;
    else slot = ((int64_t)P.ep_start[e] + m) % P.n_eps_ind;
    PTG_CHECK_INDEX(P, slot, (int64_t)P.n_eps_ind, 7);
    double v = (double)P.eps_ind[slot];
    ep_h = (int)(v * P.eps_len_d * 24);           // :60 / :491
    ep_d = (int)(v * P.eps_len_d);                // :61 / :492
}

// ---- plant transition, PTGEnv.step (:336-440) + _perform_sim_step (:525-557) ----------------------------------
// Split in two so that the memory the transition needs can be requested long before the branchy part runs:
//   plan_transition   : from (action, packed state, T flags) alone -> which handler runs, which table it binds,
//                       which argmin-LUT column it reads, whether it draws noise
//   apply_transition  : consumes the LUT value / noise draw -> new (i, j, meta) and the step-table entry
#define PTG_KIND_CONT 0      // _cont            (:559-570)
#define PTG_KIND_DRAW 1      // _standby / _cooldown / _startup (:572-625): argmin + noise
#define PTG_KIND_PARTIAL 2   // _partial         (:627-691)
#define PTG_KIND_FULL 3      // _full            (:693-756)
struct Plan {
    int kind, ds, col;       // col = argmin-LUT column or -1
};

// The part of the plan that only depends on (action, Meth_State, T flags, hot_cold): evaluated once per combination on
// the host (ptg_create -> DevParams.plan_lut, 400 x uint16) so that the step kernel replaces three divergent branches
// by one shared-memory look-up.  Packed: [0,2) kind | [2,7) table of a DRAW transition | [7,10) argmin-LUT column + 1
// (0 = none) | bit 10 hot_cold after the hysteresis (:339-342).
PTG_HD uint32_t plan_pack(int action, int state, int tflags, uint32_t hot_old) {
    const uint32_t hot = (tflags & PTG_TF_COLD) ? 0u : (tflags & PTG_TF_HOT) ? 1u : hot_old;
    // the 5x5 match (:368-440): bit (5*action + state) set <=> the step continues the current table
    const uint32_t CONT = (1u << 0) | (1u << 6) | (7u << 12) | (15u << 15) | (7u << 20) | (1u << 24);
    uint32_t kind, ds = 0, col1 = 0;
    if ((CONT >> (5 * action + state)) & 1u) {
        kind = PTG_KIND_CONT;
    } else if (action <= PTG_STARTUP) {
        kind = PTG_KIND_DRAW;
        ds = action == PTG_STANDBY ? ((tflags & PTG_TF_SBUP) ? PTG_DS_STANDBY_UP : PTG_DS_STANDBY_DOWN)        // :579-582
           : action == PTG_COOLDOWN ? PTG_DS_COOLDOWN
                                    : (hot ? PTG_DS_STARTUP_HOT : PTG_DS_STARTUP_COLD);                       // :616-619
        // LUT column of table id 0..5: startup_cold 3, startup_hot 4, cooldown 0, standby_down 2, standby_up 1, op1 5
        col1 = ((0x512043u >> (4 * ds)) & 15u) + 1u;
    } else {
        kind = action == PTG_PARTIAL_LOAD ? PTG_KIND_PARTIAL : PTG_KIND_FULL;
    }
    return kind | (ds << 2) | (col1 << 7) | (hot << 10);
}
#define PTG_PLAN_LUT_SIZE (25 * 16)
PTG_HD int plan_lut_index(int action, int state, int tflags, uint32_t hot_old) {
    return ((5 * action + state) << 4) | (tflags << 1) | (int)hot_old;
}

#if defined(__CUDACC__)
__device__ __forceinline__ Plan plan_transition(int action, uint32_t& meta, int tflags, const uint16_t* plan_lut) {
    const int state = meta & 7;
    const uint32_t ent = plan_lut[plan_lut_index(action, state, tflags, (meta >> 3) & 1u)];
    // hot/cold hysteresis (:339-342) and current_action (:347) are updated first, like the reference does
    meta = (meta & ~(0xfu << 3)) | (((ent >> 10) & 1u) << 3) | ((uint32_t)action << 4);
    Plan p;
    p.kind = ent & 3u;
    p.ds = (ent >> 2) & 31u;
    p.col = (int)((ent >> 7) & 7u) - 1;
    if (p.kind == PTG_KIND_CONT) p.ds = meta_tab(meta, state);
    else if (p.kind == PTG_KIND_PARTIAL && meta_tab(meta, PTG_FULL_LOAD) == PTG_DS_OP2_START_F) p.col = 5;   // :641
    return p;
}
#endif

// Load-change chain entry of a _partial / _full transition (:636-688, :702-754): only needs the state BEFORE the
// transition, so the step kernel requests it together with the argmin-LUT value (PTG_CHAIN_EARLY) instead of inside
// the divergent transition code, where the gather would only be issued after the noise-drawing lanes are done.
#ifndef PTG_CHAIN_EARLY
#define PTG_CHAIN_EARLY 1
#endif
__device__ __forceinline__ uint32_t chain_lookup(const DevParams& P, const Plan& p, uint32_t meta, int i, int j) {
    const bool to_partial = p.kind == PTG_KIND_PARTIAL;
    const uint32_t src = meta_tab(meta, to_partial ? PTG_FULL_LOAD : PTG_PARTIAL_LOAD);
    int chain = -1;
    if (to_partial) chain = src == PTG_DS_OP2_START_F ? PTG_CHAIN_P_FROM_OP2F : src == PTG_DS_OP3_P_F ? PTG_CHAIN_P_FROM_OP3 : -1;
    else chain = src == PTG_DS_OP1_START_P ? PTG_CHAIN_F_FROM_OP1 : src == PTG_DS_OP8_F_P ? PTG_CHAIN_F_FROM_OP8 : -1;
    uint32_t c = chain_pack(to_partial ? PTG_DS_OP8_F_P : PTG_DS_OP3_P_F, PTG_CHAIN_J_ONE, 0);   // :684-688, :750-754
    if (chain >= 0) {
        int t_op = min(i + j * P.S, P.chain_top);
        PTG_CHECK_INDEX(P, t_op, P.chain_top + 1, 3);
        c = __ldg(P.chain_tab + chain * (P.chain_top + 1) + t_op);
    }
    return c;
}

__device__ __forceinline__ int apply_transition(const DevParams& P, int64_t e, const Plan& p, int& i, int& j,
                                                uint32_t& meta, int lut_val, const uint64_t* zig_kiwi,
                                                uint32_t& draws_ep, uint32_t chain_c) {
    const int S = P.S;
    int state = meta & 7, ds = p.ds, next_state, change = 0;
    if (p.kind == PTG_KIND_CONT) {
        j += 1;
        next_state = (state == PTG_STARTUP) ? PTG_PARTIAL_LOAD : state;          // :386-388
        change = (state == PTG_STARTUP);
    } else if (p.kind == PTG_KIND_DRAW) {
        state = (meta >> 4) & 7;                                                 // new state = action
        if (state == PTG_STARTUP) {                                              // :611-614
            meta = meta_set_tab(meta_set_tab(meta, PTG_PARTIAL_LOAD, PTG_DS_OP1_START_P), PTG_FULL_LOAD, PTG_DS_OP2_START_F);
            next_state = PTG_PARTIAL_LOAD; change = 1;
        } else {
            next_state = state;
        }
        meta = meta_set_tab(meta, state, ds);
        double nz = 0.0;
        if (P.noise_mode != PTG_NOISE_OFF) nz = draw_noise(P, e, zig_kiwi, request_rng(P, e), draws_ep);
        i = jitter_index(lut_val, nz);
        j = 1;
    } else {                                                                     // _partial / _full
        const bool to_partial = p.kind == PTG_KIND_PARTIAL;
        state = next_state = to_partial ? PTG_PARTIAL_LOAD : PTG_FULL_LOAD;
        const uint32_t c = PTG_CHAIN_EARLY ? chain_c : chain_lookup(P, p, meta, i, j);
        ds = c >> 27;
        const uint32_t rule = (c >> 25) & 3u;
        const int new_i = (int)(c & 0x1ffffffu);
        if (rule == PTG_CHAIN_J_KEEP) { j += 1; }
        else if (rule == PTG_CHAIN_J_DEVELOPED) { i = P.i_fully_developed; j = P.j_fully_developed; }
        else { i = rule == PTG_CHAIN_J_ARGMIN ? lut_val : new_i; j = 1; }
        meta = meta_set_tab(meta, state, ds);
    }
    // _perform_sim_step (:525-557) reduced to index arithmetic: the window [pos-S, pos) clipped/padded at L is
    // entry min(pos - S, L) of table ds (built by k_build_step_tab with the same padding / hand-over rules)
    PTG_CHECK_INDEX(P, ds, PTG_N_DATASETS, 1);
    const int L = P.ds_len[ds];
    const long long pos = (long long)i + (long long)j * S;
    long long start = pos - S;
    if (pos >= L) {
        state = next_state;
        if (start < L) {            // time_overhead < S
            if (change) { i = (int)(pos - L); j = 0; }
        } else {
            start = L;              // S copies of the last row; (i, j) unchanged
        }
    }
    meta = (meta & ~7u) | (uint32_t)state;
    PTG_CHECK_INDEX(P, start, (long long)L + 1, 2);
    return P.ent_off[ds] + (int)start;
}

#endif  // __CUDACC__

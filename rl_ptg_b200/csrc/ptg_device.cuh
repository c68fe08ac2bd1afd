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

#define PTG_BLOCK 256            // threads per CTA of the step kernels
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
// core = {i, j, k, meta};  meta bits: [0,3) Meth_State, 3 hot_cold, 4 standby is standby_up, 5 startup is
// startup_hot, [6,11) partial table id, [11,16) full table id, [16,19) current_action
struct Meta {
    int state, hot_cold, sb_up, su_hot, part_ds, full_ds, cur_action;
};
PTG_HD Meta meta_unpack(uint32_t m) {
    Meta r;
    r.state = m & 7; r.hot_cold = (m >> 3) & 1; r.sb_up = (m >> 4) & 1; r.su_hot = (m >> 5) & 1;
    r.part_ds = (m >> 6) & 31; r.full_ds = (m >> 11) & 31; r.cur_action = (m >> 16) & 7;
    return r;
}
PTG_HD uint32_t meta_pack(const Meta& r) {
    return (uint32_t)r.state | ((uint32_t)r.hot_cold << 3) | ((uint32_t)r.sb_up << 4) | ((uint32_t)r.su_hot << 5) |
           ((uint32_t)r.part_ds << 6) | ((uint32_t)r.full_ds << 11) | ((uint32_t)r.cur_action << 16);
}

// Scalars of the reward model, shared by the table-build kernel and the eval-mode info path.
struct RewardConsts {
    double convert_mol_to_Nm3, H_u_CH4, H_u_H2, dt_water, cp_water, rho_water, Molar_mass_CO2, Molar_mass_H2O,
           h_H2O_evap, eeg_el_price, heat_price, o2_price, water_price, min_load_electrolyzer, max_h2_volumeflow,
           eta_CHP, sim_step_d, penalty;
    int32_t b_s3, _pad;
};

// --- per-env noise generator: ONE 64 B line; sector 0 is read-modify-written by a draw, sector 1 is constant ---
struct __align__(16) RngRec {
    uint64_t s_hi, s_lo;   // PCG64 state
    int64_t draws;         // normal draws consumed so far (tape position in tape mode)
    uint64_t _pad0;
    uint64_t i_hi, i_lo;   // PCG64 increment
    uint64_t _pad1, _pad2;
};
static_assert(sizeof(RngRec) == 64, "RngRec layout");

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
    int32_t raw;               // 1 = raw observation design
    int32_t continuous, eval_mode, noise_mode, schedule_mode;
    int32_t has_penalty;
    int32_t prefetch_distance;     // envs between a CTA and the CTA whose state it prefetches into L2
    int32_t action_bytes;          // element size of the action tensor of the current launch
    // tables (device)
    const StepEntry* step_tab;
    const StepMeans* mean_tab;     // same indexing as step_tab
    const int32_t* argmin_lut;     // [n_vals][6]
    const float4* hour_tab;        // [n_hours][nv]
    const DayRow* day_tab;         // [n_days]
    const ClockRow* clock_tab;     // [eps_sim_steps + 1]
    const int64_t* eps_ind;        // [n_eps_ind] or nullptr
    const double* pot0;            // e_r_b[1, 0, :] fp64 (info "Pot_Reward")
    const double* pf0;             // e_r_b[2, 0, :] fp64 (info "Part_Full")
    ZigTables zig;
    int32_t ent_off[PTG_N_DATASETS];   // first entry of table d in step_tab (table d has len+1 entries)
    int32_t ds_len[PTG_N_DATASETS];
    // reset constants (:105-138)
    int32_t reset_i, reset_tinfo;
    float reset_norm[6];
    double reset_flow[5];
    // thresholds
    int32_t time1_start_p_f, time2_start_f_p, time_p_f, time_f_p;
    int32_t time1_p_f_p, time2_p_f_p, time3_p_f_p, time34_p_f_p, time4_p_f_p, time45_p_f_p, time5_p_f_p;
    int32_t time1_f_p_f, time2_f_p_f, time23_f_p_f, time3_f_p_f, time34_f_p_f, time4_f_p_f, time45_f_p_f,
            time5_f_p_f;
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
    int32_t* tinfo;                // (T value id << 3) | flags
    int2* ep;                      // act_ep_h, act_ep_d
    double* ep_ret;                // Monitor-style running return (sum of returned rewards)
    int32_t* ep_count;             // constructor/resets consumed (m)
    int32_t* ep_start;             // SUBPROC schedule: start offset
    uint32_t* nchg;                // state changes this episode (only maintained when has_penalty)
    RngRec* rng;                   // per-env generator record (one 64 B line)
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

__device__ __forceinline__ int cur_table(const Meta& m) {
    switch (m.state) {
        case PTG_STANDBY: return m.sb_up ? PTG_DS_STANDBY_UP : PTG_DS_STANDBY_DOWN;
        case PTG_COOLDOWN: return PTG_DS_COOLDOWN;
        case PTG_STARTUP: return m.su_hot ? PTG_DS_STARTUP_HOT : PTG_DS_STARTUP_COLD;
        case PTG_PARTIAL_LOAD: return m.part_ds;
        default: return m.full_ds;
    }
}

// One draw of np_random.normal(0, noise, size=1)[0].  Out of line on purpose: only transitions into
// standby/cooldown/startup draw, and the 128-bit PCG64 arithmetic would otherwise inflate the register
// footprint of every thread of the step kernel.
__device__ __noinline__ double draw_noise(const DevParams& P, int64_t e) {
    if (P.noise_mode == PTG_NOISE_OFF) return 0.0;
    ulonglong2* rec = reinterpret_cast<ulonglong2*>(P.rng + e);
    const ulonglong2 st = rec[0], dr = rec[1];
    const int64_t d = (int64_t)dr.x;
    if (P.noise_mode == PTG_NOISE_TAPE) {
        rec[1] = make_ulonglong2((unsigned long long)(d + 1), 0ull);
        if (d >= P.tape_len) { atomicOr(P.err, PTG_EBIT_TAPE); return 0.0; }
        return P.tape[e * P.tape_len + d];
    }
    const ulonglong2 inc = rec[2];
    Pcg64 g = {st.x, st.y, inc.x, inc.y};
    const double z = pcg64_standard_normal(g, P.zig);
    rec[0] = make_ulonglong2(g.s_hi, g.s_lo);
    rec[1] = make_ulonglong2((unsigned long long)(d + 1), 0ull);
    return 0.0 + P.noise * z;      // random_normal: loc + scale * standard_normal
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

// Decode the action of env e (discrete id, or continuous Box(-1,1) -> 5 bins, :346-355)
__device__ __forceinline__ int decode_action(const DevParams& P, const void* actions, int dtype, int64_t idx,
                                             int prev_action) {
    if (!P.continuous) {
        long long a;
        if (dtype == PTG_ACT_I64) a = ((const long long*)actions)[idx];
        else if (dtype == PTG_ACT_I32) a = ((const int*)actions)[idx];
        else if (dtype == PTG_ACT_U8) a = ((const unsigned char*)actions)[idx];
        else a = (long long)((const float*)actions)[idx];
        if (a < 0 || a > 4) { atomicOr(P.err, PTG_EBIT_ACTION); a = PTG_COOLDOWN; }
        return (int)a;
    }
    double a = (double)((const float*)actions)[idx];
    int act = prev_action;                        // a >= 1.0: no interval matches, previous action is kept
#pragma unroll
    for (int ival = 5; ival >= 0; --ival)
        if (P.prob_thre[ival] > a) act = (ival + 4) % 5;   // first matching ival wins (descending scan)
    return act;
}

// episode schedule: which eps_ind entry does env (global id) use for its m-th constructor/reset
__device__ __forceinline__ void episode_offsets(const DevParams& P, int64_t e, int32_t m, int& ep_h, int& ep_d) {
    if (P.eps_ind == nullptr) { ep_h = 0; ep_d = 0; return; }     // val/test env, :63-64
    int64_t gid = P.env_id_offset + e, slot;
    if (P.schedule_mode == PTG_SCHED_DUMMY) slot = (P.n_envs_global * (int64_t)m + gid) % P.n_eps_ind;
    else slot = ((int64_t)P.ep_start[e] + m) % P.n_eps_ind;
    double v = (double)P.eps_ind[slot];
    ep_h = (int)(v * P.eps_len_d * 24);           // :60 / :491
    ep_d = (int)(v * P.eps_len_d);                // :61 / :492
}

// Which argmin-LUT column (if any) the coming transition will read: known as soon as (action, state, T flags) are,
// so the LUT gather and the RNG-state prefetch can be issued before the branchy transition code runs.
//   return: column 0..5, or -1 when no _get_index is evaluated;  draws = 1 when the transition draws noise
__device__ __forceinline__ int argmin_column(int action, const Meta& m, int tflags, int& draws) {
    const int hot = (tflags & PTG_TF_COLD) ? 0 : (tflags & PTG_TF_HOT) ? 1 : m.hot_cold;     // :339-342
    draws = 0;
    if (action == PTG_STANDBY && m.state != PTG_STANDBY) { draws = 1; return (tflags & PTG_TF_SBUP) ? 1 : 2; }
    if (action == PTG_COOLDOWN && m.state != PTG_COOLDOWN) { draws = 1; return 0; }
    if (action == PTG_STARTUP && m.state < PTG_STARTUP) { draws = 1; return hot ? 4 : 3; }
    if (action == PTG_PARTIAL_LOAD && m.state == PTG_FULL_LOAD && m.full_ds == PTG_DS_OP2_START_F) return 5;
    return -1;
}

// The plant transition of PTGEnv.step (:336-440 + _perform_sim_step) -> new (core, tinfo) and the step-table
// entry index that holds this step's window statistics.  `lut_val` = argmin_lut[vid][argmin_column(...)].
__device__ __forceinline__ int plant_transition(const DevParams& P, int64_t e, int action, int& i, int& j, Meta& m,
                                                int32_t tinfo_in, int lut_val) {
    const int tflags = tinfo_in & 7;
    if (tflags & PTG_TF_COLD) m.hot_cold = 0;          // :339-342
    else if (tflags & PTG_TF_HOT) m.hot_cold = 1;
    m.cur_action = action;
    const int S = P.S;
    int state = m.state, ds, next_state, change = 0;
    bool cont;
    switch (action) {                                   // the 5x5 match, :368-440
        case PTG_STANDBY: cont = (state == PTG_STANDBY); break;
        case PTG_COOLDOWN: cont = (state == PTG_COOLDOWN); break;
        case PTG_STARTUP: cont = (state >= PTG_STARTUP); break;
        case PTG_PARTIAL_LOAD: cont = (state != PTG_FULL_LOAD); break;
        default: cont = (state != PTG_PARTIAL_LOAD); break;
    }
    if (cont) {                                         // _cont, :559-570
        j += 1;
        ds = cur_table(m);
        next_state = (state == PTG_STARTUP) ? PTG_PARTIAL_LOAD : state;
        change = (state == PTG_STARTUP);
    } else if (action <= PTG_STARTUP) {                 // _standby / _cooldown / _startup, :572-625
        if (action == PTG_STANDBY) {
            m.sb_up = (tflags & PTG_TF_SBUP) ? 1 : 0;
            ds = m.sb_up ? PTG_DS_STANDBY_UP : PTG_DS_STANDBY_DOWN;
            next_state = PTG_STANDBY;
        } else if (action == PTG_COOLDOWN) {
            ds = PTG_DS_COOLDOWN; next_state = PTG_COOLDOWN;
        } else {
            m.part_ds = PTG_DS_OP1_START_P; m.full_ds = PTG_DS_OP2_START_F;
            m.su_hot = m.hot_cold;
            ds = m.su_hot ? PTG_DS_STARTUP_HOT : PTG_DS_STARTUP_COLD;
            next_state = PTG_PARTIAL_LOAD; change = 1;
        }
        state = action;
        i = jitter_index(lut_val, draw_noise(P, e));
        j = 1;
    } else if (action == PTG_PARTIAL_LOAD) {            // _partial, :627-691
        state = PTG_PARTIAL_LOAD; next_state = PTG_PARTIAL_LOAD;
        const int time_op = i + j * S;
        int nds = PTG_DS_OP8_F_P, ni = 0, nj = 1;
        if (m.full_ds == PTG_DS_OP2_START_F) {
            if (time_op < P.time2_start_f_p) { nds = PTG_DS_OP1_START_P; ni = lut_val; }
        } else if (m.full_ds == PTG_DS_OP3_P_F) {
            if (time_op < P.time1_p_f_p) { ni = P.i_fully_developed; nj = P.j_fully_developed; }
            else if (P.time1_p_f_p < time_op && time_op < P.time2_p_f_p) { nds = PTG_DS_OP4_P_F_P_5; ni = i; nj = j + 1; }
            else if (P.time2_p_f_p < time_op && time_op < P.time_p_f) { nds = PTG_DS_OP4_P_F_P_5; ni = P.time2_p_f_p; }
            else if (P.time_p_f < time_op && time_op < P.time34_p_f_p) { nds = PTG_DS_OP5_P_F_P_10; ni = P.time3_p_f_p; }
            else if (P.time34_p_f_p < time_op && time_op < P.time45_p_f_p) { nds = PTG_DS_OP6_P_F_P_15; ni = P.time4_p_f_p; }
            else if (P.time45_p_f_p < time_op && time_op < P.time5_p_f_p) { nds = PTG_DS_OP7_P_F_P_22; ni = P.time5_p_f_p; }
        }
        m.part_ds = nds; ds = nds; i = ni; j = nj;
    } else {                                            // _full, :693-756
        state = PTG_FULL_LOAD; next_state = PTG_FULL_LOAD;
        const int time_op = i + j * S;
        int nds = PTG_DS_OP3_P_F, ni = 0, nj = 1;
        if (m.part_ds == PTG_DS_OP1_START_P) {
            if (time_op < P.time1_start_p_f) nds = PTG_DS_OP2_START_F;
        } else if (m.part_ds == PTG_DS_OP8_F_P) {
            if (time_op < P.time1_f_p_f) { ni = P.i_fully_developed; nj = P.j_fully_developed; }
            else if (P.time1_f_p_f < time_op && time_op < P.time_f_p) { nds = PTG_DS_OP9_F_P_F_5; ni = i; nj = j + 1; }
            else if (P.time_f_p < time_op && time_op < P.time23_f_p_f) { nds = PTG_DS_OP9_F_P_F_5; ni = P.time2_f_p_f; }
            else if (P.time23_f_p_f < time_op && time_op < P.time34_f_p_f) { nds = PTG_DS_OP10_F_P_F_10; ni = P.time3_f_p_f; }
            else if (P.time34_f_p_f < time_op && time_op < P.time45_f_p_f) { nds = PTG_DS_OP11_F_P_F_15; ni = P.time4_f_p_f; }
            else if (P.time45_f_p_f < time_op && time_op < P.time5_f_p_f) { nds = PTG_DS_OP12_F_P_F_20; ni = P.time5_f_p_f; }
        }
        m.full_ds = nds; ds = nds; i = ni; j = nj;
    }
    // _perform_sim_step, :525-557, reduced to index arithmetic: the window [pos-S, pos) clipped/padded at L is
    // entry min(pos - S, L) of table ds (built by k_build_step_tab with the same padding / hand-over rules)
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
    m.state = state;
    return P.ent_off[ds] + (int)start;
}

#endif  // __CUDACC__

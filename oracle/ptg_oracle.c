/*
 * ptg_oracle.c -- CPU restatement of RL_PtG's PTGEnv (TEST INFRASTRUCTURE, not product code).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * It deliberately follows the reference's algorithm literally -- it slices the real S-row window out of the
 * experimental tables, runs a real argmin over the whole temperature column, and sums with numpy's pairwise
 * scheme -- so that it shares NO look-up tables or shortcuts with the CUDA product path it checks.
 *
 * Parity pinning: tests/test_oracle_golden.py checks this file against tests/golden/ vectors recorded from the
 * UNMODIFIED reference env (tests/golden/gen_golden.py, run where /root/reference is mounted).
 *
 * Citations are file:line in the reference checkout (env/ptg_gym_env.py unless stated).
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/ptg_b200.h"

#define NCOL 7

typedef struct OracleEnv {
    /* persistent plant state */
    int32_t meth_state, i, j, k, hot_cold;
    int32_t standby_ds, startup_ds, partial_ds, full_ds;   /* which table self.standby/.startup/.partial/.full is */
    int32_t current_action;
    int32_t act_ep_h, act_ep_d;
    int32_t episode_count;
    int32_t state_change;
    int64_t draws;
    double t_cat;
    double h2, ch4, h2res, h2o, heat;     /* Meth_*_flow, Meth_el_heating */
    double cum_rew;                       /* :330 (without penalty) */
    double ep_return;                     /* Monitor: sum of returned rewards */
    /* current market slices (e_r_b_act / g_e_act) */
    int64_t t_hour, t_day;
    double sin_h, cos_h;
    /* reward constituents for info (:266-275) */
    double ch4_rev, steam_rev, o2_rev, eua_rev, chp_rev, heat_cost, ely_cost, water_cost, rew;
} OracleEnv;

typedef struct OracleCtx {
    PtgConfig cfg;
    const double* op[PTG_N_DATASETS];
    int64_t op_rows[PTG_N_DATASETS];
    const double* e_r_b; int64_t n_hours;
    const double* g_e; int64_t n_days;
    const int64_t* eps_ind; int64_t n_eps_ind;
    int64_t ep_index;          /* the reference's module-global ep_index (:9) */
    int32_t step_size;         /* :66 */
    int32_t b_s3;              /* :76-77 */
    double prob_thre[6];       /* :151-155 */
    int64_t n_envs;
    OracleEnv* envs;
    int32_t obs_dim;
    double* win;               /* scratch [step_size][7] per thread is allocated by callers */
} OracleCtx;

/* ---- numpy's pairwise summation (np.add.reduce over a strided fp64 vector), so np.average matches bitwise ---- */
static double np_pairwise_sum(const double* a, int64_t n, int64_t stride) {
    if (n < 8) {
        double res = 0.0;
        for (int64_t i = 0; i < n; ++i) res += a[i * stride];
        return res;
    } else if (n <= 128) {
        double r[8];
        for (int q = 0; q < 8; ++q) r[q] = a[q * stride];
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int q = 0; q < 8; ++q) r[q] += a[(i + q) * stride];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i * stride];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum(a, n2, stride) + np_pairwise_sum(a + n2 * stride, n - n2, stride);
    }
}

/* _get_index (:514-523): first index of the minimum of |T - t_cat| */
static int64_t get_index(const OracleCtx* c, int ds, double t_cat) {
    const double* tab = c->op[ds];
    int64_t n = c->op_rows[ds], best = 0;
    double bestv = fabs(tab[1] - t_cat);
    for (int64_t r = 1; r < n; ++r) {
        double v = fabs(tab[r * NCOL + 1] - t_cat);
        if (v < bestv) { bestv = v; best = r; }
    }
    return best;
}

static inline double e_r_b_at(const OracleCtx* c, int typ, int a, int64_t t) {
    return c->e_r_b[((int64_t)typ * c->cfg.price_ahead + a) * c->n_hours + t];
}
static inline double g_e_at(const OracleCtx* c, int typ, int a, int64_t d) {
    return c->g_e[((int64_t)typ * 2 + a) * c->n_days + d];
}

/*
 * _perform_sim_step (:525-557).  Materialises op_range into `win` ([step_size][7], possibly fewer rows when a
 * next_operation is shorter than the overhead -- numpy slicing semantics) and returns the row count.
 */
static int64_t perform_sim_step(const OracleCtx* c, double* win, int operation, int initial_state,
                                int next_operation, int next_state, int64_t* idx, int64_t* j, int change,
                                int* r_state) {
    const int64_t S = c->step_size, total = c->op_rows[operation];
    const double* tab = c->op[operation];
    int64_t n = 0;
    if (*idx + *j * S < total) {
        *r_state = initial_state;
        int64_t lo = *idx + (*j - 1) * S, hi = *idx + *j * S;
        for (int64_t r = lo; r < hi; ++r) memcpy(win + (n++) * NCOL, tab + r * NCOL, NCOL * sizeof(double));
    } else {
        *r_state = next_state;
        int64_t overhead = (*idx + *j * S) - total;
        if (overhead < S) {
            int64_t lo = *idx + (*j - 1) * S;
            if (lo > total) lo = total;
            for (int64_t r = lo; r < total; ++r) memcpy(win + (n++) * NCOL, tab + r * NCOL, NCOL * sizeof(double));
            if (change) {
                *idx = overhead;
                *j = 0;
                int64_t avail = c->op_rows[next_operation] < overhead ? c->op_rows[next_operation] : overhead;
                for (int64_t r = 0; r < avail; ++r)
                    memcpy(win + (n++) * NCOL, c->op[next_operation] + r * NCOL, NCOL * sizeof(double));
            } else {
                for (int64_t r = 0; r < overhead; ++r)
                    memcpy(win + (n++) * NCOL, tab + (total - 1) * NCOL, NCOL * sizeof(double));
            }
        } else {
            for (int64_t r = 0; r < S; ++r) memcpy(win + (n++) * NCOL, tab + (total - 1) * NCOL, NCOL * sizeof(double));
        }
    }
    return n;
}

static double next_noise(const OracleCtx* c, OracleEnv* e, const double* tape, int64_t tape_len, int64_t env_idx,
                         int* err) {
    if (c->cfg.noise_mode == PTG_NOISE_OFF) return 0.0;
    if (e->draws >= tape_len) { *err = PTG_ERR_NOISE_TAPE; return 0.0; }
    return tape[env_idx * tape_len + (e->draws++)];
}

/* i = int(max(argmin + normal, 0))  (:584-585) */
static int64_t jitter_index(int64_t idx, double nz) {
    double v = (double)idx + nz;
    if (0.0 > v) v = 0.0;          /* Python max(v, 0): 0 only when 0 > v */
    return (int64_t)v;             /* int(): truncation toward zero */
}

static void refresh_market(const OracleCtx* c, OracleEnv* e, int64_t h_step, int64_t d_step, double clock_hours) {
    e->t_hour = e->act_ep_h + h_step;    /* :446 */
    e->t_day = e->act_ep_d + d_step;     /* :447 */
    e->sin_h = sin(2 * M_PI * clock_hours);   /* :449-450 */
    e->cos_h = cos(2 * M_PI * clock_hours);
}

/* _initialize_op_rew (:105-138) + the episode bookkeeping of __init__/reset (:59-68, :490-497) */
static void env_init(OracleCtx* c, OracleEnv* e) {
    if (c->eps_ind) {
        double v = (double)c->eps_ind[c->ep_index % c->n_eps_ind];
        e->act_ep_h = (int32_t)(v * c->cfg.eps_len_d * 24);
        e->act_ep_d = (int32_t)(v * c->cfg.eps_len_d);
        c->ep_index++;
    } else {
        e->act_ep_h = 0; e->act_ep_d = 0;
    }
    e->episode_count++;
    refresh_market(c, e, 0, 0, 0.0);
    e->meth_state = PTG_COOLDOWN;
    e->standby_ds = PTG_DS_STANDBY_DOWN;
    e->startup_ds = PTG_DS_STARTUP_COLD;
    e->partial_ds = PTG_DS_OP1_START_P;
    e->full_ds = PTG_DS_OP2_START_F;
    e->t_cat = 16;
    e->i = (int32_t)get_index(c, PTG_DS_COOLDOWN, e->t_cat);
    e->j = 0;
    const double* row = c->op[PTG_DS_COOLDOWN] + (int64_t)e->i * NCOL;
    e->h2 = row[2]; e->ch4 = row[3]; e->h2res = row[4]; e->h2o = row[5]; e->heat = row[6];
    e->hot_cold = 0;
    e->state_change = 0;
    e->ch4_rev = e->steam_rev = e->o2_rev = e->eua_rev = e->chp_rev = 0.0;
    e->heat_cost = e->ely_cost = e->water_cost = e->rew = 0.0;
    e->cum_rew = 0.0;
    e->ep_return = 0.0;
    e->k = 0;
}

/* _normalize_observations + _get_obs (:206-249); key order as in the dicts at :222-249 */
static void write_obs(const OracleCtx* c, const OracleEnv* e, double* o) {
    const PtgConfig* g = &c->cfg;
    int pa = g->price_ahead, p = 0;
    if (g->raw_modified == 0) {
        for (int a = 0; a < pa; ++a) o[p++] = (e_r_b_at(c, 0, a, e->t_hour) - g->el_l_b) / (g->el_u_b - g->el_l_b);
        for (int a = 0; a < 2; ++a) o[p++] = (g_e_at(c, 0, a, e->t_day) - g->gas_l_b) / (g->gas_u_b - g->gas_l_b);
        for (int a = 0; a < 2; ++a) o[p++] = (g_e_at(c, 1, a, e->t_day) - g->eua_l_b) / (g->eua_u_b - g->eua_l_b);
    } else {
        for (int a = 0; a < pa; ++a) o[p++] = (e_r_b_at(c, 1, a, e->t_hour) - g->rew_l_b) / (g->rew_u_b - g->rew_l_b);
        for (int a = 0; a < pa; ++a) o[p++] = e_r_b_at(c, 2, a, e->t_hour);
    }
    o[p++] = (double)e->meth_state;
    o[p++] = (e->t_cat - g->T_l_b) / (g->T_u_b - g->T_l_b);
    o[p++] = (e->h2 - g->h2_l_b) / (g->h2_u_b - g->h2_l_b);
    o[p++] = (e->ch4 - g->ch4_l_b) / (g->ch4_u_b - g->ch4_l_b);
    o[p++] = (e->h2res - g->h2_res_l_b) / (g->h2_res_u_b - g->h2_res_l_b);
    o[p++] = (e->h2o - g->h2o_l_b) / (g->h2o_u_b - g->h2o_l_b);
    o[p++] = (e->heat - g->heat_l_b) / (g->heat_u_b - g->heat_l_b);
    o[p++] = e->sin_h;
    o[p++] = e->cos_h;
}

/* _get_info (:251-278), row-major [24] */
static void write_info(const OracleCtx* c, const OracleEnv* e, double* f) {
    f[0] = e->k;
    f[1] = e_r_b_at(c, 0, 0, e->t_hour);
    f[2] = g_e_at(c, 0, 0, e->t_day);
    f[3] = g_e_at(c, 1, 0, e->t_day);
    f[4] = e->meth_state;
    f[5] = e->current_action;
    f[6] = e->hot_cold;
    f[7] = e->t_cat;
    f[8] = e->h2; f[9] = e->ch4; f[10] = e->h2o; f[11] = e->heat;
    f[12] = e->ch4_rev; f[13] = e->steam_rev; f[14] = e->o2_rev; f[15] = e->eua_rev; f[16] = e->chp_rev;
    f[17] = -e->heat_cost; f[18] = -e->ely_cost; f[19] = -e->water_cost;
    f[20] = e->rew;
    f[21] = e->cum_rew;
    f[22] = e_r_b_at(c, 1, 0, e->t_hour);
    f[23] = e_r_b_at(c, 2, 0, e->t_hour);
}

/* _get_reward (:280-334) */
static double get_reward(const OracleCtx* c, OracleEnv* e) {
    const PtgConfig* g = &c->cfg;
    double el = e_r_b_at(c, 0, 0, e->t_hour), gas = g_e_at(c, 0, 0, e->t_day), eua = g_e_at(c, 1, 0, e->t_day);
    double ch4_volumeflow = e->ch4 * g->convert_mol_to_Nm3;
    double h2_res_volumeflow = e->h2res * g->convert_mol_to_Nm3;
    double Q_ch4 = ch4_volumeflow * g->H_u_CH4 * 1000;
    double Q_h2_res = h2_res_volumeflow * g->H_u_H2 * 1000;
    e->ch4_rev = (Q_ch4 + Q_h2_res) * gas;
    double power_chp = Q_ch4 * g->eta_CHP * c->b_s3;
    double Q_chp = Q_ch4 * (1 - g->eta_CHP) * c->b_s3;
    e->chp_rev = power_chp * g->eeg_el_price;
    double Q_steam = e->h2o * (g->dt_water * g->cp_water + g->h_H2O_evap) / 3600;
    e->steam_rev = (Q_steam + Q_chp) * g->heat_price;
    double h2_volumeflow = e->h2 * g->convert_mol_to_Nm3;
    double o2_volumeflow = 1.0 / 2 * h2_volumeflow * 3600;
    e->o2_rev = o2_volumeflow * g->o2_price;
    double co2 = e->ch4 * g->Molar_mass_CO2 / 1000;
    e->eua_rev = co2 / 1000 * 3600 * eua * 100;
    e->heat_cost = e->heat / 1000 * el;
    double load = h2_volumeflow / g->max_h2_volumeflow, eta;
    if (load < g->min_load_electrolyzer) eta = 0.02;
    else eta = (0.598 - 0.325 * pow(load, 2) + 0.218 * pow(load, 3) + 0.01 * pow(load, -1)
                - 1.68 * pow(10, -3) * pow(load, -2) + 2.51 * pow(10, -5) * pow(load, -3));
    e->ely_cost = h2_volumeflow * g->H_u_H2 * 1000 / eta * el;
    double elec_costs = e->heat_cost + e->ely_cost;
    double water_elec = e->h2 * g->Molar_mass_H2O / 1000 * 3600;
    e->water_cost = (e->h2o + water_elec) / g->rho_water * g->water_price;
    e->rew = (e->ch4_rev + e->chp_rev + e->steam_rev + e->eua_rev + e->o2_rev - elec_costs - e->water_cost)
             * g->sim_step / 3600;
    e->cum_rew += e->rew;
    if (e->state_change) e->rew -= g->reward_level * g->state_change_penalty;
    return e->rew;
}

/* _partial (:627-691): choose the partial-load table from the previous full-load table and time_op */
static void choose_partial(const OracleCtx* c, OracleEnv* e, int64_t* i, int64_t* j) {
    const PtgConfig* g = &c->cfg;
    int64_t time_op = *i + *j * c->step_size;
    int ds = PTG_DS_OP8_F_P; int64_t ni = 0, nj = 1;
    if (e->full_ds == PTG_DS_OP2_START_F) {
        if (time_op < g->time2_start_f_p) { ds = PTG_DS_OP1_START_P; ni = get_index(c, ds, e->t_cat); nj = 1; }
    } else if (e->full_ds == PTG_DS_OP3_P_F) {
        if (time_op < g->time1_p_f_p) { ds = PTG_DS_OP8_F_P; ni = g->i_fully_developed; nj = g->j_fully_developed; }
        else if (g->time1_p_f_p < time_op && time_op < g->time2_p_f_p) { ds = PTG_DS_OP4_P_F_P_5; ni = *i; nj = *j + 1; }
        else if (g->time2_p_f_p < time_op && time_op < g->time_p_f) { ds = PTG_DS_OP4_P_F_P_5; ni = g->time2_p_f_p; }
        else if (g->time_p_f < time_op && time_op < g->time34_p_f_p) { ds = PTG_DS_OP5_P_F_P_10; ni = g->time3_p_f_p; }
        else if (g->time34_p_f_p < time_op && time_op < g->time45_p_f_p) { ds = PTG_DS_OP6_P_F_P_15; ni = g->time4_p_f_p; }
        else if (g->time45_p_f_p < time_op && time_op < g->time5_p_f_p) { ds = PTG_DS_OP7_P_F_P_22; ni = g->time5_p_f_p; }
    }
    e->partial_ds = ds; *i = ni; *j = nj;
}

/* _full (:693-756) */
static void choose_full(const OracleCtx* c, OracleEnv* e, int64_t* i, int64_t* j) {
    const PtgConfig* g = &c->cfg;
    int64_t time_op = *i + *j * c->step_size;
    int ds = PTG_DS_OP3_P_F; int64_t ni = 0, nj = 1;
    if (e->partial_ds == PTG_DS_OP1_START_P) {
        if (time_op < g->time1_start_p_f) ds = PTG_DS_OP2_START_F;
    } else if (e->partial_ds == PTG_DS_OP8_F_P) {
        if (time_op < g->time1_f_p_f) { ds = PTG_DS_OP3_P_F; ni = g->i_fully_developed; nj = g->j_fully_developed; }
        else if (g->time1_f_p_f < time_op && time_op < g->time_f_p) { ds = PTG_DS_OP9_F_P_F_5; ni = *i; nj = *j + 1; }
        else if (g->time_f_p < time_op && time_op < g->time23_f_p_f) { ds = PTG_DS_OP9_F_P_F_5; ni = g->time2_f_p_f; }
        else if (g->time23_f_p_f < time_op && time_op < g->time34_f_p_f) { ds = PTG_DS_OP10_F_P_F_10; ni = g->time3_f_p_f; }
        else if (g->time34_f_p_f < time_op && time_op < g->time45_f_p_f) { ds = PTG_DS_OP11_F_P_F_15; ni = g->time4_f_p_f; }
        else if (g->time45_f_p_f < time_op && time_op < g->time5_f_p_f) { ds = PTG_DS_OP12_F_P_F_20; ni = g->time5_f_p_f; }
    }
    e->full_ds = ds; *i = ni; *j = nj;
}

/* PTGEnv.step (:336-481).  Returns terminated. */
static int env_step(OracleCtx* c, OracleEnv* e, int64_t env_idx, const void* actions, int action_dtype,
                    const double* tape, int64_t tape_len, double* win, double* reward, int* err) {
    const PtgConfig* g = &c->cfg;
    int k = e->k;
    if (e->t_cat <= g->t_cat_startup_cold) e->hot_cold = 0;          /* :339-342 */
    else if (e->t_cat >= g->t_cat_startup_hot) e->hot_cold = 1;
    int previous_state = e->meth_state;

    if (g->action_type == 0) {                                        /* :346-347 */
        int64_t a;
        if (action_dtype == PTG_ACT_I64) a = ((const int64_t*)actions)[env_idx];
        else if (action_dtype == PTG_ACT_I32) a = ((const int32_t*)actions)[env_idx];
        else a = ((const uint8_t*)actions)[env_idx];
        if (a < 0 || a > 4) { *err = PTG_ERR_INVALID_ACTION; a = 1; }
        e->current_action = (int32_t)a;
    } else {                                                          /* :348-355 */
        double a = (double)((const float*)actions)[env_idx];
        for (int ival = 0; ival < 6; ++ival)
            if (c->prob_thre[ival] > a) { e->current_action = (ival - 1 + 5) % 5; break; }   /* actions[-1] wraps */
    }

    int action = e->current_action, state = e->meth_state;
    int64_t i = e->i, j = e->j;
    int ds, next_ds, next_state, change = 0, r_state;
    /* the 5x5 match (:368-440); "cont" = _cont (:559-570) */
    int cont;
    switch (action) {
        case PTG_STANDBY: cont = (state == PTG_STANDBY); break;
        case PTG_COOLDOWN: cont = (state == PTG_COOLDOWN); break;
        case PTG_STARTUP: cont = (state == PTG_STARTUP || state == PTG_PARTIAL_LOAD || state == PTG_FULL_LOAD); break;
        case PTG_PARTIAL_LOAD: cont = (state != PTG_FULL_LOAD); break;
        default: cont = (state != PTG_PARTIAL_LOAD); break;
    }
    if (cont) {
        j += 1;
        switch (state) {
            case PTG_STANDBY: ds = next_ds = e->standby_ds; next_state = state; break;
            case PTG_COOLDOWN: ds = next_ds = PTG_DS_COOLDOWN; next_state = state; break;
            case PTG_STARTUP: ds = e->startup_ds; next_ds = e->partial_ds; next_state = PTG_PARTIAL_LOAD; change = 1; break;
            case PTG_PARTIAL_LOAD: ds = next_ds = e->partial_ds; next_state = PTG_PARTIAL_LOAD; break;
            default: ds = next_ds = e->full_ds; next_state = PTG_FULL_LOAD; break;
        }
    } else if (action == PTG_STANDBY) {                               /* _standby :572-589 */
        state = PTG_STANDBY;
        e->standby_ds = (e->t_cat <= g->t_cat_standby) ? PTG_DS_STANDBY_UP : PTG_DS_STANDBY_DOWN;
        ds = next_ds = e->standby_ds; next_state = state;
        i = jitter_index(get_index(c, ds, e->t_cat), next_noise(c, e, tape, tape_len, env_idx, err)); j = 1;
    } else if (action == PTG_COOLDOWN) {                              /* _cooldown :591-603 */
        state = PTG_COOLDOWN;
        ds = next_ds = PTG_DS_COOLDOWN; next_state = state;
        i = jitter_index(get_index(c, ds, e->t_cat), next_noise(c, e, tape, tape_len, env_idx, err)); j = 1;
    } else if (action == PTG_STARTUP) {                               /* _startup :605-625 */
        state = PTG_STARTUP;
        e->partial_ds = PTG_DS_OP1_START_P; e->full_ds = PTG_DS_OP2_START_F;
        e->startup_ds = (e->hot_cold == 0) ? PTG_DS_STARTUP_COLD : PTG_DS_STARTUP_HOT;
        ds = e->startup_ds; next_ds = e->partial_ds; next_state = PTG_PARTIAL_LOAD; change = 1;
        i = jitter_index(get_index(c, ds, e->t_cat), next_noise(c, e, tape, tape_len, env_idx, err)); j = 1;
    } else if (action == PTG_PARTIAL_LOAD) {                          /* _partial */
        state = PTG_PARTIAL_LOAD;
        choose_partial(c, e, &i, &j);
        ds = next_ds = e->partial_ds; next_state = PTG_PARTIAL_LOAD;
    } else {                                                          /* _full */
        state = PTG_FULL_LOAD;
        choose_full(c, e, &i, &j);
        ds = next_ds = e->full_ds; next_state = PTG_FULL_LOAD;
    }
    int64_t n = perform_sim_step(c, win, ds, state, next_ds, next_state, &i, &j, change, &r_state);
    e->meth_state = r_state; e->i = (int32_t)i; e->j = (int32_t)j;

    double clock_hours = (double)((int64_t)(k + 1) * g->sim_step) / 3600;    /* :442 */
    double clock_days = clock_hours / 24;
    refresh_market(c, e, (int64_t)floor(clock_hours), (int64_t)floor(clock_days), clock_hours);
    if (e->t_hour >= c->n_hours || e->t_day >= c->n_days) {
        *err = PTG_ERR_DATA_RANGE;
        if (e->t_hour >= c->n_hours) e->t_hour = c->n_hours - 1;
        if (e->t_day >= c->n_days) e->t_day = c->n_days - 1;
    }

    e->t_cat = win[(n - 1) * NCOL + 1];                                       /* :452 */
    e->h2 = np_pairwise_sum(win + 2, n, NCOL) / (double)n;                    /* :454-458 np.average */
    e->ch4 = np_pairwise_sum(win + 3, n, NCOL) / (double)n;
    e->h2res = np_pairwise_sum(win + 4, n, NCOL) / (double)n;
    e->h2o = np_pairwise_sum(win + 5, n, NCOL) / (double)n;
    e->heat = np_pairwise_sum(win + 6, n, NCOL) / (double)n;

    e->state_change = (previous_state != e->meth_state);                      /* :463-466 */
    *reward = get_reward(c, e);
    e->ep_return += *reward;
    int terminated = (e->k == g->eps_sim_steps - 6);                          /* :508-511 */
    return terminated;
}

/* ------------------------------------------------------------------------------------------------------ */
/* exported API (ctypes)                                                                                    */
/* ------------------------------------------------------------------------------------------------------ */
int ptg_oracle_obs_dim(const PtgConfig* cfg) {
    return cfg->raw_modified == 0 ? cfg->price_ahead + 4 + 9 : 2 * cfg->price_ahead + 9;
}

OracleCtx* ptg_oracle_create(const PtgConfig* cfg, const PtgTables* t, int64_t n_envs) {
    OracleCtx* c = (OracleCtx*)calloc(1, sizeof(OracleCtx));
    c->cfg = *cfg;
    for (int d = 0; d < PTG_N_DATASETS; ++d) { c->op[d] = t->op[d]; c->op_rows[d] = t->op_rows[d]; }
    c->e_r_b = t->e_r_b; c->n_hours = t->n_hours;
    c->g_e = t->g_e; c->n_days = t->n_days;
    c->eps_ind = t->eps_ind; c->n_eps_ind = t->n_eps_ind;
    c->step_size = (int32_t)(cfg->sim_step / cfg->time_step_op);
    c->b_s3 = cfg->scenario == 3 ? 1 : 0;
    double prob_ival = (1.0 - (-1.0)) / 5;
    for (int q = 0; q < 6; ++q) c->prob_thre[q] = -1 + q * prob_ival;
    c->n_envs = n_envs;
    c->obs_dim = ptg_oracle_obs_dim(cfg);
    c->envs = (OracleEnv*)calloc((size_t)n_envs, sizeof(OracleEnv));
    c->ep_index = 0;
    for (int64_t e = 0; e < n_envs; ++e) {       /* constructors, in env order (make_vec_env) */
        c->envs[e].current_action = PTG_COOLDOWN;   /* :143 */
        c->envs[e].episode_count = -1;
        env_init(c, &c->envs[e]);
    }
    return c;
}

void ptg_oracle_destroy(OracleCtx* c) {
    if (!c) return;
    free(c->envs);
    free(c);
}

/* VecEnv.reset(): env order; obs row-major [n_envs][obs_dim]; info row-major [n_envs][24] or NULL */
int ptg_oracle_reset(OracleCtx* c, const uint8_t* mask, double* obs, double* info) {
    for (int64_t e = 0; e < c->n_envs; ++e) {
        if (mask && !mask[e]) continue;
        env_init(c, &c->envs[e]);
        if (obs) write_obs(c, &c->envs[e], obs + e * c->obs_dim);
        if (info) write_info(c, &c->envs[e], info + e * PTG_N_INFO);
    }
    return 0;
}

/*
 * VecEnv.step_wait() with DummyVecEnv auto-reset.  All outputs row-major per env; terminal_obs / info may be
 * NULL.  `threads` > 1 steps disjoint env ranges on pthreads (envs are independent); the auto-reset pass
 * itself always runs serially in env order so that the global ep_index is consumed exactly like DummyVecEnv
 * does.
 */
typedef struct StepJob {
    OracleCtx* c; const void* actions; int action_dtype; const double* tape; int64_t tape_len;
    double *obs, *reward, *info; uint8_t* done; int64_t lo, hi; int err;
} StepJob;

static void* step_range(void* arg) {
    StepJob* jb = (StepJob*)arg;
    OracleCtx* c = jb->c;
    double* win = (double*)malloc((size_t)c->step_size * NCOL * sizeof(double));
    for (int64_t e = jb->lo; e < jb->hi; ++e) {
        OracleEnv* env = &c->envs[e];
        int err = 0;
        double r = 0.0;
        int term = env_step(c, env, e, jb->actions, jb->action_dtype, jb->tape, jb->tape_len, win, &r, &err);
        if (jb->obs) write_obs(c, env, jb->obs + e * c->obs_dim);
        if (jb->info) write_info(c, env, jb->info + e * PTG_N_INFO);   /* uses k before the increment (:254) */
        env->k += 1;                                                   /* :476 */
        if (jb->reward) jb->reward[e] = r;
        if (jb->done) jb->done[e] = (uint8_t)term;
        if (err < jb->err) jb->err = err;
    }
    free(win);
    return NULL;
}

int ptg_oracle_step(OracleCtx* c, const void* actions, int action_dtype, const double* tape, int64_t tape_len,
                    double* obs, double* reward, uint8_t* done, double* terminal_obs, double* info,
                    double* episode_return, int32_t* episode_length, int auto_reset, int threads) {
    int err_all = 0;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if ((int64_t)threads > c->n_envs) threads = (int)c->n_envs;
    StepJob jobs[256];
    pthread_t tid[256];
    int64_t per = (c->n_envs + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        StepJob jb = {c, actions, action_dtype, tape, tape_len, obs, reward, info, done, t * per,
                      (t + 1) * per < c->n_envs ? (t + 1) * per : c->n_envs, 0};
        jobs[t] = jb;
        if (t > 0) pthread_create(&tid[t], NULL, step_range, &jobs[t]);
    }
    step_range(&jobs[0]);
    for (int t = 1; t < threads; ++t) pthread_join(tid[t], NULL);
    for (int t = 0; t < threads; ++t) if (jobs[t].err < err_all) err_all = jobs[t].err;
    if (auto_reset) {
        for (int64_t e = 0; e < c->n_envs; ++e) {
            if (!done[e]) continue;
            OracleEnv* env = &c->envs[e];
            if (terminal_obs && obs) memcpy(terminal_obs + e * c->obs_dim, obs + e * c->obs_dim, c->obs_dim * sizeof(double));
            if (episode_return) episode_return[e] = env->ep_return;
            if (episode_length) episode_length[e] = env->k;
            env_init(c, env);
            if (obs) write_obs(c, env, obs + e * c->obs_dim);
        }
    }
    return err_all;
}

/* state snapshot, SoA host arrays as in PtgStateSoA */
int ptg_oracle_get_state(const OracleCtx* c, const PtgStateSoA* s) {
    for (int64_t e = 0; e < c->n_envs; ++e) {
        const OracleEnv* v = &c->envs[e];
        if (s->meth_state) s->meth_state[e] = v->meth_state;
        if (s->i) s->i[e] = v->i;
        if (s->j) s->j[e] = v->j;
        if (s->k) s->k[e] = v->k;
        if (s->hot_cold) s->hot_cold[e] = v->hot_cold;
        if (s->standby_ds) s->standby_ds[e] = v->standby_ds;
        if (s->startup_ds) s->startup_ds[e] = v->startup_ds;
        if (s->partial_ds) s->partial_ds[e] = v->partial_ds;
        if (s->full_ds) s->full_ds[e] = v->full_ds;
        if (s->current_action) s->current_action[e] = v->current_action;
        if (s->act_ep_h) s->act_ep_h[e] = v->act_ep_h;
        if (s->act_ep_d) s->act_ep_d[e] = v->act_ep_d;
        if (s->episode_count) s->episode_count[e] = v->episode_count;
        if (s->draws) s->draws[e] = v->draws;
        if (s->t_cat) s->t_cat[e] = v->t_cat;
        if (s->cum_reward) s->cum_reward[e] = v->ep_return;
    }
    return 0;
}

/* numpy's sum for tests of the device window-mean kernel */
double ptg_oracle_pairwise_sum(const double* a, int64_t n, int64_t stride) { return np_pairwise_sum(a, n, stride); }
int64_t ptg_oracle_get_index(const OracleCtx* c, int ds, double t_cat) { return get_index(c, ds, t_cat); }

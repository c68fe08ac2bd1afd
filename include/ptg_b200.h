/*
 * ptg_b200.h -- C ABI of the B200-native batched PtG environment (libptg_b200.so).
 *
 * Drop-in boundary for the PTGEnv.step()/reset() hot path of SimMarkt/RL_PtG.  The reference has no FFI of
 * its own (it is pure Python); every entry point below names the reference interface it replaces
 * (file:line relative to the reference checkout).  Plain pointers and sizes only -- no torch types.
 *
 * Conventions
 *   - return value 0 = PTG_OK, negative = PtgStatus error; ptg_last_error() gives the message (thread-local).
 *     The reference signals these conditions with Python `assert` (env/ptg_gym_env.py:46,158,204,357,440).
 *   - "host" pointers are read during the call only.  "device" pointers are caller-owned CUDA memory
 *     (e.g. torch tensors' data_ptr()); the library never frees them and keeps none past the call.
 *   - every launch goes to the `stream` argument (a cudaStream_t passed as void*); no call synchronises the
 *     device except ptg_create / ptg_destroy / ptg_get_state / ptg_set_state / ptg_poll_error.
 *   - a handle is bound to one device and is not thread-safe.  ptg_step / ptg_step_many / ptg_episode_stats /
 *     ptg_allreduce_stats / the train-side entry points launch on `stream` without switching the calling thread's
 *     current device: they return PTG_ERR_INVALID_ARGUMENT when that device is not the handle's (one process per GPU
 *     is the intended arrangement); ptg_create / ptg_reset / get / set_state select the handle's device themselves.
 */
#ifndef PTG_B200_H_
#define PTG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTG_ABI_VERSION 3   /* 2: PtgConfig.obs_layout, PtgIO.windows_changed, train-side entry points
                               3: PtgStateSoA.rng / .state_changes, PtgConfig.no_auto_reset, windows_changed carries the step serial,
                                  ptg_allreduce_stats + ptg_nccl_* (cross-rank statistics inside the library) */
#define PTG_N_DATASETS 17
#define PTG_N_INFO 24        /* fields of PTGEnv._get_info(), env/ptg_gym_env.py:251-278 */
#define PTG_MAX_PRICE_AHEAD 16

/* Methanation time-series tables, in the order of the reference's op_data_files (src/rl_utils.py:104-110). */
enum PtgDataset {
    PTG_DS_STARTUP_COLD = 0, PTG_DS_STARTUP_HOT = 1, PTG_DS_COOLDOWN = 2, PTG_DS_STANDBY_DOWN = 3,
    PTG_DS_STANDBY_UP = 4, PTG_DS_OP1_START_P = 5, PTG_DS_OP2_START_F = 6, PTG_DS_OP3_P_F = 7,
    PTG_DS_OP4_P_F_P_5 = 8, PTG_DS_OP5_P_F_P_10 = 9, PTG_DS_OP6_P_F_P_15 = 10, PTG_DS_OP7_P_F_P_22 = 11,
    PTG_DS_OP8_F_P = 12, PTG_DS_OP9_F_P_F_5 = 13, PTG_DS_OP10_F_P_F_10 = 14, PTG_DS_OP11_F_P_F_15 = 15,
    PTG_DS_OP12_F_P_F_20 = 16
};

/* Plant states == action ids (config_env.yaml:73-78; env/ptg_gym_env.py:50-56,142). */
enum PtgPlantState { PTG_STANDBY = 0, PTG_COOLDOWN = 1, PTG_STARTUP = 2, PTG_PARTIAL_LOAD = 3, PTG_FULL_LOAD = 4 };

typedef enum PtgStatus {
    PTG_OK = 0,
    PTG_ERR_INVALID_ARGUMENT = -1,   /* bad config value (reference: assert at ptg_gym_env.py:46,158,204) */
    PTG_ERR_CUDA = -2,               /* CUDA runtime error */
    PTG_ERR_UNSUPPORTED = -3,        /* valid in the reference but outside this build's limits */
    PTG_ERR_INVALID_ACTION = -4,     /* device saw an action outside 0..4 (reference: ptg_gym_env.py:347,440) */
    PTG_ERR_DATA_RANGE = -5,         /* device indexed past the market tables (reference: IndexError) */
    PTG_ERR_NOISE_TAPE = -6,         /* tape-mode noise ran past the end of the tape */
    PTG_ERR_NCCL = -7                /* libnccl.so.2 could not be loaded, or an NCCL call failed (ptg_nccl_*, ptg_allreduce_stats) */
} PtgStatus;

enum PtgActionDtype { PTG_ACT_I64 = 0, PTG_ACT_I32 = 1, PTG_ACT_U8 = 2, PTG_ACT_F32 = 3 };

/* Source of the Gaussian row jitter drawn on transitions into standby/cooldown/startup
 * (np_random.normal(0, noise, size=1)[0]; env/ptg_gym_env.py:584-585,598-599,620-621). */
enum PtgNoiseMode {
    PTG_NOISE_NUMPY = 0,   /* on-device PCG64 + ziggurat: bit-identical to numpy's Generator(PCG64(seed)).normal */
    PTG_NOISE_TAPE = 1,    /* caller-supplied pre-drawn fp64 tape, consumed in order (parity harness) */
    PTG_NOISE_OFF = 2      /* noise term is exactly 0.0 (no draw) */
};

/* Episode schedule: which eps_ind entry env e uses for its m-th constructor/reset (m = 0 is the constructor).
 * Reference: module-global ep_index, env/ptg_gym_env.py:9,37-44,59-62,487-493. */
enum PtgScheduleMode {
    PTG_SCHED_DUMMY = 0,   /* DummyVecEnv order: eps_ind[(n_envs_global*m + global_env_id) mod len] */
    PTG_SCHED_SUBPROC = 1  /* SubprocVecEnv: per-env start offset in [0, n_eps_loops), then +1 per reset */
};

/* Layout of the observation buffers (PtgIO.obs / terminal_obs).
 *   PTG_OBS_KEY_MAJOR : one contiguous [n_envs, key_dim] block per observation key (the Dict of
 *                       env/ptg_gym_env.py:222-249 as zero-copy views); METH_STATUS holds int32 bit patterns.
 *   PTG_OBS_FLAT      : one [ptg_features_dim()]-float row per env in the column order of SB3's CombinedExtractor
 *                       (see ptg_features): what a MultiInputPolicy consumes, written by the step kernel itself.
 *                       Built for price_ahead == 13 and train_or_eval = train (else PTG_ERR_UNSUPPORTED). */
enum PtgObsLayout { PTG_OBS_KEY_MAJOR = 0, PTG_OBS_FLAT = 1 };

/* All scalar knobs the env reads from its constructor dict (src/rl_utils.py:345-403). */
typedef struct PtgConfig {
    int32_t abi_version;            /* = PTG_ABI_VERSION */
    int32_t scenario;               /* 1, 2, 3 -> b_s3 = (scenario == 3), ptg_gym_env.py:76-77 */
    int32_t raw_modified;           /* 0 = "raw", 1 = "mod" */
    int32_t action_type;            /* 0 = "discrete", 1 = "continuous" */
    int32_t train_or_eval;          /* 0 = "train" (step info empty), 1 = "eval" (24-field info every step) */
    int32_t price_ahead;
    int32_t sim_step;               /* [s] */
    int32_t time_step_op;           /* [s] */
    int32_t eps_sim_steps;
    int32_t noise_mode;             /* PtgNoiseMode */
    int32_t schedule_mode;          /* PtgScheduleMode */
    int32_t n_eps_loops;
    /* load-change thresholds [rows], config_env.yaml:133-155 */
    int32_t time1_start_p_f, time2_start_f_p, time_p_f, time_f_p;
    int32_t time1_p_f_p, time2_p_f_p, time23_p_f_p, time3_p_f_p, time34_p_f_p, time4_p_f_p, time45_p_f_p,
            time5_p_f_p;
    int32_t time1_f_p_f, time2_f_p_f, time23_f_p_f, time3_f_p_f, time34_f_p_f, time4_f_p_f, time45_f_p_f,
            time5_f_p_f;
    int32_t i_fully_developed, j_fully_developed;
    int32_t obs_layout;             /* PtgObsLayout: 0 = key-major blocks (default), 1 = flat feature rows */
    int32_t no_auto_reset;          /* 0 (default) = SB3 VecEnv semantics: a done env is reset inside the step (the
                                       returned obs is the reset obs).  1 = Gymnasium single-env semantics
                                       (env/ptg_gym_env.py:476-481): the returned obs is the terminal obs and the env
                                       stays as it is until ptg_reset -- which alone consumes the next eps_ind entry */
    double noise;                   /* sigma [rows] */
    double eps_len_d;               /* [d] */
    double state_change_penalty;
    double reward_level;            /* r_0 = reward_level[0] */
    double convert_mol_to_Nm3, H_u_CH4, H_u_H2, dt_water, cp_water, rho_water, Molar_mass_CO2, Molar_mass_H2O,
           h_H2O_evap, eeg_el_price, heat_price, o2_price, water_price, min_load_electrolyzer,
           max_h2_volumeflow, eta_CHP;
    double t_cat_standby, t_cat_startup_cold, t_cat_startup_hot;
    double el_l_b, el_u_b, gas_l_b, gas_u_b, eua_l_b, eua_u_b, T_l_b, T_u_b, h2_l_b, h2_u_b, ch4_l_b, ch4_u_b,
           h2_res_l_b, h2_res_u_b, h2o_l_b, h2o_u_b, heat_l_b, heat_u_b, rew_l_b, rew_u_b;
} PtgConfig;

/* Host arrays of the constructor dict (read during ptg_create only). */
typedef struct PtgTables {
    const double* op[PTG_N_DATASETS];     /* row-major [op_rows[d]][7]: t, T_cat, n_h2, n_ch4, n_h2_res, m_h2o, P_el */
    int64_t op_rows[PTG_N_DATASETS];
    const double* e_r_b;                  /* [3][price_ahead][n_hours]  el_price, pot_rew, part_full */
    int64_t n_hours;
    const double* g_e;                    /* [2][2][n_days]  gas, eua x (today, tomorrow) */
    int64_t n_days;
    const int64_t* eps_ind;               /* training episode order, or NULL for val/test envs */
    int64_t n_eps_ind;
} PtgTables;

/* Device output/input buffers of one step()/reset() call.  obs is ONE fp32 buffer of obs_dim*n_envs elements,
 * key-major (each observation key is a contiguous [n_envs, key_dim] block, keys in the order of
 * env/ptg_gym_env.py:222-249); METH_STATUS holds int32 bit patterns.  See ptg_obs_layout(). */
typedef struct PtgIO {
    float* obs;              /* [obs_dim * n_envs] */
    float* reward;           /* [n_envs]; NULL allowed for ptg_reset */
    uint8_t* done;           /* [n_envs]; NULL allowed for ptg_reset */
    float* terminal_obs;     /* same layout as obs, written only for envs with done=1 (SB3 "terminal_observation");
                                NULL = do not record */
    double* info;            /* [PTG_N_INFO][n_envs] feature-major fp64, or NULL.  Written by ptg_reset always
                                (reference reset() returns the full info even in train mode, :503-506) and by
                                ptg_step when train_or_eval = eval.  For done envs it holds the terminal step's
                                info (SB3 keeps the pre-reset info). Meth_Action is the action id 0..4. */
    double* episode_return;  /* [n_envs] Monitor-style sum of returned rewards of the episode that just ended */
    int32_t* episode_length; /* [n_envs] its length; both written only where done=1; NULL allowed */
    uint32_t* windows_changed; /* one word or NULL (ptg_step, key-major layout): when the step changed any env's
                                  market-window blocks (Pot_Reward / Part_Full / Elec_Price / Gas_Price / EUA_Price)
                                  the kernel stores the step's serial number here (ptg_last_step_serial() right after
                                  the call; never 0).  The blocks only move when an env's clock crosses an hour or its
                                  episode ends, i.e. on one step in 3600 / sim_step.  The library never clears the
                                  word and the caller does not have to: a host mirror that finds a value other than
                                  the serial of the step it just issued may skip the transfer of those blocks (3/4 of
                                  the observation bytes). */
    uint8_t* status_u8;      /* [n_envs] or NULL (ptg_step): METH_STATUS of the returned observation as one byte per env
                                -- what a host mirror transfers instead of the 4-byte block of the obs buffer */
} PtgIO;

/* One observation key of the obs buffer. */
typedef struct PtgObsKey {
    char name[24];
    int32_t dim;             /* values per env */
    int32_t is_int32;        /* 1 for METH_STATUS */
    int64_t offset;          /* key-major: element offset of the [n_envs, dim] block inside obs;
                                flat: first column of the key inside a row (METH_STATUS: 6 one-hot floats) */
} PtgObsKey;

/* Episode statistics of finished episodes since the last clear (per rank; combine across ranks with
 * ptg_stats_combine or any all-gather of this 64-byte struct). */
typedef struct PtgEpisodeStats {
    double count, sum_return, sum_return_sq, sum_length, min_return, max_return;
    double total_steps;      /* env-steps executed since the last clear */
    double _reserved;
} PtgEpisodeStats;

typedef struct PtgHandle PtgHandle;

/* PTGEnv.__init__ for n_envs environments (env/ptg_gym_env.py:28-79) + DummyVecEnv/make_vec_env construction
 * (src/rl_utils.py:448-453).  Uploads the tables, builds the device look-up tables with CUDA kernels and
 * constructs the envs (each consumes one eps_ind entry like the reference constructor, :59-62).
 * env_id_offset / n_envs_global describe this shard of a larger env population (multi-GPU): global env id =
 * env_id_offset + local index; schedules and seeds depend on the global id only. */
int ptg_create(const PtgConfig* cfg, const PtgTables* tables, int64_t n_envs, int64_t env_id_offset,
               int64_t n_envs_global, int device, PtgHandle** out);

void ptg_destroy(PtgHandle* h);

/* VecEnv.seed(seed) + VecEnv.reset() (SB3) over PTGEnv.reset(seed) (env/ptg_gym_env.py:483-506).
 *   seeds : host int64[n_envs] or NULL.  Env e with seeds[e] >= 0 re-creates its generator as
 *           Generator(PCG64(SeedSequence(seeds[e]))) (gymnasium Env.reset(seed=...)); negative = keep stream.
 *   mask  : host uint8[n_envs] or NULL (= all).  Only masked envs are reset.
 * Writes obs (and info when io->info != NULL) for the reset envs. */
int ptg_reset(PtgHandle* h, const int64_t* seeds, const uint8_t* mask, const PtgIO* io, void* stream);

/* VecEnv.step_wait() over PTGEnv.step(action) (env/ptg_gym_env.py:336-481) with SB3 auto-reset:
 * where done=1 the returned obs is the reset obs, io->terminal_obs gets the last obs.
 *   actions : device pointer, n_envs elements of `action_dtype` (F32 = continuous Box(-1,1) actions). */
int ptg_step(PtgHandle* h, const void* actions, int action_dtype, const PtgIO* io, void* stream);

/* T consecutive steps in one launch with the per-env state kept in registers.
 *   actions : device [T][n_envs];  io buffers are [T] x the single-step shapes (obs: [T][obs_dim*n_envs], ...).
 * terminal_obs / info / episode_* are not recorded by this entry point (must be NULL). */
int ptg_step_many(PtgHandle* h, const void* actions, int action_dtype, int32_t T, const PtgIO* io, void* stream);

/* Tape-mode noise: device fp64 [n_envs][tape_len], values as returned by normal(0, noise) (already scaled). */
int ptg_set_noise_tape(PtgHandle* h, const double* tape_dev, int64_t tape_len);

/* Plant state snapshot (host SoA arrays of n_envs each; any pointer may be NULL in ptg_get_state, none in
 * ptg_set_state).  Used by parity tests and checkpointing (the reference never checkpoints env state): a restored
 * env continues bit-identically, noise stream included. */
typedef struct PtgStateSoA {
    int32_t* meth_state;     /* Meth_State */
    int32_t* i;              /* row index i */
    int32_t* j;              /* step counter j */
    int32_t* k;              /* step in episode */
    int32_t* hot_cold;
    int32_t* standby_ds;     /* PtgDataset currently bound to self.standby / startup / partial / full */
    int32_t* startup_ds;
    int32_t* partial_ds;
    int32_t* full_ds;
    int32_t* current_action;
    int32_t* act_ep_h;
    int32_t* act_ep_d;
    int32_t* episode_count;  /* m: constructor/resets consumed so far */
    int64_t* draws;          /* noise values consumed so far (tape position) */
    double* t_cat;           /* Meth_T_cat */
    double* cum_reward;      /* Monitor-style running return (== cum_rew when state_change_penalty == 0) */
    uint64_t* rng;           /* [n_envs][4]: the env's PCG64 generator {state_hi, state_lo, inc_hi, inc_lo} (np_random) */
    uint32_t* state_changes; /* Meth_State changes this episode (info cum_reward = cum_reward + penalty * this) */
} PtgStateSoA;
int ptg_get_state(PtgHandle* h, const PtgStateSoA* out);
int ptg_set_state(PtgHandle* h, const PtgStateSoA* in);

/* Deterministic reduction in one launch (warp shuffles -> per-block partials -> the last block folds them in index order) of the finished-episode
 * accumulators into *stats_dev (device, 64 bytes); clear != 0 zeroes the accumulators afterwards. */
int ptg_episode_stats(PtgHandle* h, PtgEpisodeStats* stats_dev, int clear, void* stream);

/* Host-side combine of per-rank stats (sum / min / max), e.g. after an NCCL all-gather of the 64-byte structs. */
void ptg_stats_combine(const PtgEpisodeStats* per_rank, int n_ranks, PtgEpisodeStats* out);

/* Cross-rank reduction of the episode statistics inside the library (the path's ONLY collective):
 * ncclAllGather of the 64-byte record over `nccl_comm` (an ncclComm_t as void*; one process per GPU, NVLink/NVSwitch)
 * followed by a fixed-rank-order combine kernel, everything on `stream`, no host synchronisation.  On return (in
 * stream order) *stats_dev -- this rank's record from ptg_episode_stats -- holds the combined record on every rank.
 * Replaces what SB3's Monitor/logger do with per-env episode records in one process (rollout/ep_rew_mean). */
int ptg_allreduce_stats(PtgHandle* h, void* nccl_comm, PtgEpisodeStats* stats_dev, void* stream);

/* NCCL plumbing for binders that have no communicator of their own.  libnccl.so.2 is dlopen()ed on first use (the one
 * already in the process if any -- e.g. the torch wheel's -- so there are never two NCCLs in one process).
 *   ptg_nccl_unique_id   : rank 0 creates the 128-byte id and ships it to the others by any side channel
 *   ptg_nccl_comm_create : every rank, with the handle's device current; *comm_out is an ncclComm_t
 *   ptg_nccl_comm_destroy */
#define PTG_NCCL_UNIQUE_ID_BYTES 128
int ptg_nccl_unique_id(char id_out[PTG_NCCL_UNIQUE_ID_BYTES]);
int ptg_nccl_comm_create(const char id[PTG_NCCL_UNIQUE_ID_BYTES], int n_ranks, int rank, void** comm_out);
int ptg_nccl_comm_destroy(void* comm);

/* ---- callers either side of the path; all pointers are device memory unless noted ----------- */

/* SB3 VecNormalize(venv, norm_obs=False) as the reference wraps every env (src/rl_utils.py:453, :491; defaults
 * norm_reward=True, clip_reward=10, gamma=0.99, epsilon=1e-8), split so that the batch moments can be exchanged
 * between ranks (one 24-byte all-gather) before they are folded into the running statistics:
 *   ptg_vecnorm_moments : returns = returns * gamma + reward (VecNormalize._update_reward); moments_out[3] =
 *                         {count, mean, sum of squared deviations} of the new returns (deterministic tree; one
 *                         launch; st_in = the current statistics, whose mean is the summation pivot)
 *   ptg_vecnorm_apply   : RunningMeanStd.update_from_moments with `n_batch` such records (training != 0), then
 *                         reward_out = clip(reward_in / sqrt(var + epsilon), +-clip_reward); returns[done] = 0.
 *                         st_in / st_out: {mean, var, count, pad} fp64, distinct buffers (ping-pong). */
int ptg_vecnorm_moments(PtgHandle* h, const float* reward, double* returns, double gamma, const double* st_in,
                        double* moments_out, void* stream);
int ptg_vecnorm_apply(PtgHandle* h, const float* reward_in, const uint8_t* done, double* returns, const double* st_in,
                      double* st_out, const double* moments, int32_t n_batch, int32_t training, double epsilon,
                      double clip_reward, float* reward_out, void* stream);

/* The flat feature rows SB3's MultiInputPolicy (CombinedExtractor) builds from the Dict observation: keys in the
 * sorted order of the gymnasium Dict space, METH_STATUS one-hot(6).  feat: [n_envs][ptg_features_dim()] fp32.
 *   mod: CH4_syn, Elec_Heating, H2O_DE, H2_in, H2_res, METH_STATUS[6], Part_Full[pa], Pot_Reward[pa], T_CAT, cos, sin
 *   raw: CH4_syn, EUA_Price[2], Elec_Heating, Elec_Price[pa], Gas_Price[2], H2O_DE, H2_in, H2_res, METH_STATUS[6],
 *        T_CAT, cos, sin */
int ptg_features_dim(const PtgHandle* h);
int ptg_features(PtgHandle* h, const float* obs, float* feat, void* stream);

/* SB3 RolloutBuffer.compute_returns_and_advantage (GAE(lambda)) over [T][n_envs] roll-out buffers, in numpy's
 * evaluation order and dtypes (float32 buffers, float64 carry).  PPO hyper-parameters of the reference:
 * config/config_agent.yaml:46,53 (gamma 0.973, gae_lambda 0.8002). */
int ptg_gae(int64_t n_envs, int32_t T, const float* rewards, const float* values, const uint8_t* episode_starts,
            const float* last_values, const uint8_t* last_dones, double gamma, double gae_lambda, float* advantages,
            float* returns, void* stream);

/* calculate_optimum (src/rl_opt.py:26-152) on the device: the (n_hours, 24) statistics of the theoretical optimum
 * T-OPT (columns = EnvConfiguration.stats_names, src/rl_config_env.py:44-49), whose columns 20 / 23 are the series
 * behind the Pot_Reward / Part_Full observations.  The caller evaluates the level-dependent constants once per load
 * level (off, partial, full; two evaluations of the PEM efficiency polynomial) in the reference's operation order;
 * el / gas / eua and stats_out are device pointers, stats_out is row-major [n_hours][24] fp64. */
typedef struct PtgOptLevel {
    double q_gas;        /* (Q_ch4 + Q_h2_res)                       -> ch4_revenues = q_gas * gas */
    double chp_rev, steam_rev, o2_rev;
    double k_eua;        /* Meth_CO2_mass_flow / 1000 * 3600         -> eua_revenues = k_eua * eua * 100 */
    double k_heat;       /* Meth_el_heating / 1000                   -> elec_costs_heating = k_heat * el */
    double k_ely;        /* h2_volumeflow * H_u_H2 * 1000 / eta      -> elec_costs_electrolyzer = k_ely * el */
    double water_cost;
    double stat8[8];     /* Meth_State, Meth_Action, Meth_Hot_Cold, Meth_T_cat, H2, CH4, H2O, el_heating (cols 4..11) */
} PtgOptLevel;
int ptg_calculate_optimum(const double* el, int64_t n_hours, const double* gas, const double* eua, int64_t n_days,
                          const PtgOptLevel* levels3, double* stats_out, void* stream);

/* Sticky device-side error word (invalid action, market-table overrun, tape overrun).  Synchronises `stream`.
 * Returns PTG_OK or the first error seen since the last poll and clears it. */
int ptg_poll_error(PtgHandle* h, void* stream);

/* Introspection. */
int ptg_obs_dim(const PtgHandle* h);                               /* fp32 elements per env */
int ptg_obs_layout(const PtgHandle* h, PtgObsKey* keys, int max_keys);   /* returns number of keys */
int64_t ptg_obs_elems(const PtgHandle* h);                       /* fp32 elements of one obs buffer (padded) */
int64_t ptg_num_envs(const PtgHandle* h);
int64_t ptg_bytes_per_env_step(const PtgHandle* h, int action_dtype);    /* algorithmic HBM bytes, DESIGN.md */
int ptg_kernel_launches(const PtgHandle* h, int64_t* out);               /* kernels launched by this handle */
/* Do all envs of the handle share one episode clock?  True from ptg_create / an unmasked ptg_reset on, as long as every
 * env is stepped by the same calls (always the case) and no masked reset or ptg_set_state intervened -- tracked on the
 * host, no device work.  Returns 1 and the clock encodings of the latest returned observation (Temp_hour_enc_sin / _cos,
 * env/ptg_gym_env.py:449-450, the same fp32 values the kernels write) or 0.  A host mirror can then skip the transfer of
 * those two blocks.  (Steps replayed from a captured CUDA graph do not pass through ptg_step and are not seen by this
 * tracking: after a replay treat the clock as unknown until the next unmasked ptg_reset.) */
int ptg_clock_uniform(const PtgHandle* h, float* sin_out, float* cos_out);
uint32_t ptg_last_step_serial(const PtgHandle* h);                       /* serial of the latest ptg_step (see PtgIO.windows_changed) */
/* Host twins of the on-device generator (no GPU needed): the first n values of
 * Generator(PCG64(SeedSequence(seed))).standard_normal() -- bit-identical to numpy; used by the CPU test-suite. */
int ptg_host_standard_normal(uint64_t seed, int64_t n, double* out);
/* PCG64 state after seeding: out[0..3] = state_hi, state_lo, inc_hi, inc_lo. */
int ptg_host_seed_state(uint64_t seed, uint64_t* out4);
/* Diagnostic, no reference counterpart: dependent-load latency of this GPU's L2 (out4[0], ns per hop over a 16 MB
 * ring) and DRAM (out4[1], 256 MB ring) and the SM clock a lone thread is given (out4[2], MHz).  The step kernel is bound
 * by latency x occupancy, so bench.py reports these beside its numbers (config.box_probe). */
int ptg_probe_box(int device, double* out4);
const char* ptg_last_error(void);
int ptg_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PTG_B200_H_ */

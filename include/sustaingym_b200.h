/*
 * sustaingym_b200 — C ABI of the B200-native batched step() engine.
 *
 * The reference (chrisyeh96/sustaingym @ eac2a4d) is pure Python and has no FFI layer; its
 * boundary is the Gymnasium `Env.step()/reset()` and PettingZoo `ParallelEnv` API.  The Python
 * package `sustaingym_b200` mirrors that API and calls the entry points below through ctypes.
 * Every entry point cites the reference function whose per-step work it replaces.
 *
 * Conventions
 *   - All `const T*` / `T*` arguments marked [device] are CUDA device pointers owned by the
 *     caller (torch tensors in the Python host code); the library never allocates or frees.
 *     Arguments marked [host] are host pointers (pinned memory recommended).
 *   - Batched arrays are env-major, C-contiguous: `action[B, n]`, `obs[B, obs_dim]`.
 *   - `stream` is a `cudaStream_t` passed as `void*` (0 = legacy default stream).  Calls are
 *     asynchronous on that stream unless the name ends in `_host` (those synchronise the stream
 *     before returning because they write host memory).
 *   - Return value: 0 on success, non-zero on error; `sg_last_error()` returns a message
 *     (thread-local).  Numerical non-convergence is NEVER an error: it is reported per env in a
 *     status array (this replaces `pdb.set_trace()` in sustaingym/envs/utils.py:21-24).
 *   - The library is re-entrant per stream; it keeps no global mutable state.
 */
#ifndef SUSTAINGYM_B200_H
#define SUSTAINGYM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SG_ABI_VERSION 2

int         sg_abi_version(void);
const char* sg_last_error(void);
/* Number of SMs of the current device (grid sizing helper for callers / tests). */
int         sg_device_sm_count(void);

/* ------------------------------------------------------------------------------------------
 * BuildingEnv  —  replaces sustaingym/envs/building/env.py:205-312 (step), :314-386 (reset),
 *                 :411-434 (_calc_occupower); constants from :196-203 (A_d, BD_d).
 * ---------------------------------------------------------------------------------------- */
typedef struct sg_building_model {
    int32_t n;             /* zones                                                         */
    int32_t T;             /* rows of the weather tables (env.py:141 length_of_weather)      */
    int32_t episode_len;   /* env.py:138                                                     */
    int32_t discrete;      /* 0: action is f32 [B,n] in [-ac,ac]; 1: int64 MultiDiscrete     */
    double  q_rate;        /* (1-beta)*24, env.py:179                                        */
    double  error_rate;    /* beta, env.py:180                                               */
    double  pnorm;         /* reward_pnorm (2 = Euclidean fast path)                         */
    const double* A_d;     /* [device] [n, n]    env.py:202                                  */
    const double* BD_d;    /* [device] [n, n+4]  env.py:203, columns [occ|ground,out,a..,ghi]*/
    const double* target;  /* [device] [n]                                                   */
    const double* ac_map;  /* [device] [n]                                                   */
    const double* weather; /* [device] [T, 4]: out_temp, ground_temp, ghi, metabolism per row */
} sg_building_model;

/* on done: runs <- 0, zone temperatures <- target, weather cursor keeps running (equivalent to
 * reset(seed) with the seed that maps to the current cursor, env.py:339-378).  The observation
 * returned for a finished env is then the first observation of its next episode. */
#define SG_BUILDING_FLAG_AUTORESET 16
/* n >= 16, continuous actions, pnorm 2 run on the tensor-core (DMMA) cluster kernel; this flag forces the
 * thread-per-env kernel instead (tests compare the two). */
#define SG_BUILDING_FLAG_NO_MMA 32

/* reset(): env.py:346-378 with the weather cursor already resolved by the host
 * (seed -> epoch mapping env.py:339-345 is host logic).  `T_initial` may be NULL (= target). */
int sg_building_reset(const sg_building_model* m, int64_t num_envs,
                      const int32_t* epoch_in /* [device] [B] */,
                      const double* T_initial /* [device] [B,n] or NULL */,
                      float* state /* [device] [B,n+4] out */,
                      int32_t* epoch /* [device] [B] out */, int32_t* runs /* [device] [B] out */,
                      void* stream);

/* One step() of B independent envs.  `state` is the previous observation and is overwritten
 * with the new one (the reference's state IS its f32 observation, env.py:286-296).
 * (n in {3, 5, 6, 10} and num_envs >= 512: a persistent kernel walks over 128-env tiles; the environment variable
 * SG_BLD_NO_PERSISTENT selects the one-tile-per-CTA kernel instead — same results, for comparison.) */
int sg_building_step(const sg_building_model* m, int64_t num_envs,
                     const void* action /* [device] [B,n] f32 or i64 */,
                     float* state /* [device] [B,n+4] in/out */,
                     int32_t* epoch /* [device] [B] in/out */, int32_t* runs /* [device] [B] in/out */,
                     double* reward /* [device] [B] out */, uint8_t* done /* [device] [B] out */,
                     int32_t flags /* SG_BUILDING_FLAG_* */, double* stats /* [device] [16] or NULL */,
                     void* stream);

/* `steps` consecutive step() calls fused in one launch (state stays in registers between
 * steps).  action: [steps, B, n]; obs_out: [steps, B, n+4] or NULL (only final state kept);
 * reward: [steps, B]; done: [steps, B]. */
int sg_building_rollout(const sg_building_model* m, int64_t num_envs, int32_t steps,
                        const void* action, float* state, int32_t* epoch, int32_t* runs,
                        float* obs_out, double* reward, uint8_t* done,
                        int32_t flags, double* stats, void* stream);

/* Host-buffer variant of sg_building_step: copies `h_action` to `d_action`, steps (same `flags` / `d_stats`
 * as sg_building_step), copies state/reward/done back, synchronises. d_* are caller-owned device
 * scratch of the same shapes. */
int sg_building_step_host(const sg_building_model* m, int64_t num_envs,
                          const void* h_action, float* h_obs, double* h_reward, uint8_t* h_done,
                          void* d_action, float* d_state, int32_t* d_epoch, int32_t* d_runs,
                          double* d_reward, uint8_t* d_done, int32_t flags, double* d_stats, void* stream);

/* ------------------------------------------------------------------------------------------
 * EVChargingEnv  —  replaces sustaingym/envs/evcharging/env.py:229-291 (step), :200-221 +
 *   :178-198 + :473-500 + envs/utils.py:6-24 (action projection, cvxpy+MOSEK), :340-379
 *   (_to_schedule), :381-394 (_get_observation), :431-464 (_get_reward), and the un-vendored
 *   acnportal.acnsim Simulator.step / ChargingNetwork / Linear2StageBattery inner loop
 *   (call sites env.py:283-284, 385, 445, 450-451).
 * ---------------------------------------------------------------------------------------- */
#define SG_EV_MAX_STATIONS    64
#define SG_EV_MAX_CLASSES     8
#define SG_EV_MAX_CONSTRAINTS 16
#define SG_EV_MOER_COLS       37   /* load_moer.py:364-377: col 0 historical + 36 forecasts  */
#define SG_EV_MOER_ROWS       289
#define SG_EV_EPISODE_STEPS   288
#define SG_EV_TABLE_BYTES     4096

/* One charging session = one row of the events DataFrame (event_generation.py:125-147) after
 * the requested-energy cap (:169-170).  Slot k of station i on day d lives at
 * sessions[(d*K + k)*n + i]; slots of a station are ordered by arrival; unused slots have
 * arrival = departure = 32767. */
typedef struct sg_ev_session {
    int16_t arrival;        /* period index 0..287                                            */
    int16_t departure;
    int16_t est_departure;
    int16_t reserved;
    double  energy;         /* requested kWh                                                   */
} sg_ev_session;

typedef struct sg_ev_model {
    int32_t n_stations;      /* 54 (Caltech)                                                 */
    int32_t n_classes;       /* distinct columns of A~ (env.py:485-486); 5 for Caltech        */
    int32_t n_constraints;   /* rows of constraint_matrix; 8 for Caltech                      */
    int32_t moer_steps;      /* k = moer_forecast_steps (env.py:117), 1..36                   */
    int32_t max_sessions;    /* K: session slots per station per day in the trace tables      */
    int32_t n_days;          /* D: day traces resident on the device                          */
    int32_t n_moer_days;     /* Dm: MOER day tables resident on the device                    */
    int32_t project;         /* project_action_in_env (env.py:118)                            */
    const uint8_t* station_class; /* [device] [n] class id of each station                    */
    const uint8_t* station_type;  /* [device] [n] 0: {0,6..32} A (min_pilot 6), 1: {0,8,..,32}*/
    const double*  con_re;        /* [device] [16, 8] Re of the per-class constraint coefficient,
                                     zero-padded to [SG_EV_MAX_CONSTRAINTS, SG_EV_MAX_CLASSES]  */
    const double*  con_im;        /* [device] [16, 8] Im, same layout                          */
    const double*  magnitudes;    /* [device] [16] limits in A, zero-padded                    */
    const void*    sessions;      /* [device] [D, K, n] 16-byte records (sg_ev_session)        */
    const float*   moer;          /* [device] [Dm, 289, 37]                                   */
    const uint32_t* arrivals;     /* [device] [D, max_day_sessions + 1] the day's PluginEvents (event_generation.py:188-196)
                                     sorted by (arrival, station): arrival | station << 16 | slot << 24,
                                     terminated by 0x7fff entries                                       */
    int32_t max_day_sessions;     /* row length of `arrivals` minus the terminator                  */
    int32_t reserved;
    const void* tables;           /* [device] SG_EV_TABLE_BYTES, 16-byte aligned: derived constant tables filled
                                     once per model by sg_ev_prepare_tables (the kernels stage them in shared memory) */
    const float* obs_tail;        /* [device] [Dm, 289, SG_EV_TAIL_FLOATS(k)] or NULL: the last k+2 entries of the
                                     observation for every (MOER day, t) — forecasted_moer[k] | prev_moer | t/288 —
                                     laid out as they appear in the row (filled by sg_ev_prepare_obs_tail), so the step
                                     kernel copies them with asynchronous 8-byte copies instead of 39 scalar loads;
                                     NULL: the kernel reads `moer` directly                                        */
} sg_ev_model;
#define SG_EV_TAIL_FLOATS(k) (((k) + 3) & ~1)   /* k + 2 rounded up to an even count */

/* Per-env state.
 *   slots : a plugged-in session is one 32-byte record
 *             { f64 requested kWh, f64 delivered kWh, i32 departure, i32 est_departure, i32 arrival, i32 trace slot }
 *           Each env owns a block of n records holding its live sessions as a PACKED LIST in station order: the
 *           record of station i is entry popcount(occupancy & ((1 << i) - 1)); the list is updated in place
 *           (sg_ev_export_state gives dense per-station views).
 *   header: one 64-byte record per env, two 32-byte halves
 *             { i32 t (0..288, env.py:168), i32 day (trace index in [0,D)), i32 moer_day (MOER table index
 *               in [0,Dm)), i32 arrival cursor (sessions of the day with arrival <= t),
 *               u64 occupancy mask (bit i: station i has a plugged-in session), 8 bytes padding }
 *             { f64[3] cumulative profit, carbon_cost, excess_charge (env.py:460-462), 8 bytes padding } */
typedef struct sg_ev_state {
    void* slots;          /* [device] [B, n] 32-byte session records, 32-byte aligned */
    void* header;         /* [device] [B] 64-byte env records, 64-byte aligned        */
} sg_ev_state;

/* status bits written per env by sg_ev_step */
#define SG_EV_STATUS_PROJECTED   1  /* clipped action violated the network -> solver ran      */
#define SG_EV_STATUS_MAXITER     2  /* solver hit its iteration cap (result still feasible)   */

/* flags */
#define SG_EV_ACTION_F32      0
#define SG_EV_ACTION_I64      1   /* DiscreteActionWrapper input (wrappers.py:43-45)          */
#define SG_EV_FLAG_AUTORESET  16  /* on done: advance day (mod D), reset state, obs = reset obs */
#define SG_EV_FLAG_STATIC_TILES 32 /* sg_ev_step: grid-stride tile assignment instead of the dynamic scheduler
                                     (diagnostic; results are identical)                           */

/* `work` scratch of sg_ev_step: SG_EV_WORK_INTS(B) int32, zero-initialised once by the caller; every call leaves
 * it zeroed again.  Three 128-byte lines of counters — [0] entries appended to the exact-projection work list, [1]
 * entries claimed by the helper warps; [32] tiles handed out by the dynamic tile scheduler; [64] tile warps that
 * left the tile loop, [65] helper warps that finished — then [96...] the work list (env index + 1). */
#define SG_EV_WORK_HEADER_INTS 96
#define SG_EV_WORK_INTS(B)     ((B) + SG_EV_WORK_HEADER_INTS)

/* Fills the derived constant tables of a model (f32 copies of the constraint coefficients, squared
 * limits, the pilot -> kW table, padded class / EVSE-type maps) into `tables`; call once per model,
 * then store the pointer in m->tables.  Replaces EVChargingEnv.__init__ / _init_action_projection
 * bookkeeping (env.py:131-134, 178-198). */
int sg_ev_prepare_tables(const sg_ev_model* m, void* tables /* [device] [SG_EV_TABLE_BYTES] */, void* stream);

/* Fills the observation-tail table of a model from m->moer (env.py:389-392: prev_moer = moer[t,0],
 * forecasted_moer = moer[t,1:k+1], timestep = t/288); store the pointer in m->obs_tail afterwards. */
int sg_ev_prepare_obs_tail(const sg_ev_model* m, float* obs_tail /* [device] [Dm, 289, SG_EV_TAIL_FLOATS(k)] */, void* stream);

/* reset(): env.py:293-338 for envs whose `reset_mask` byte is non-zero (NULL = all).
 * `day_in`/`moer_day_in` pick the trace/MOER table (the seed->day map is host logic,
 * event_generation.py:273-282).  Writes the reset observation. */
int sg_ev_reset(const sg_ev_model* m, const sg_ev_state* s, int64_t num_envs,
                const uint8_t* reset_mask, const int32_t* day_in, const int32_t* moer_day_in,
                float* obs /* [device] [B, 2n+k+2] */, void* stream);

/* One step() of B independent envs.
 *   action : [B, n] f32 in [0,1] (or i64 in [0,bins) with SG_EV_ACTION_I64)
 *   obs    : [B, 2n+k+2] f32, gymnasium flatten order (sorted keys):
 *            demands[n] | est_departures[n] | forecasted_moer[k] | prev_moer[1] | timestep[1]
 *   reward : [B] f64;  done: [B] u8 (terminated; truncated is always False, env.py:291)
 *   pilots : [B, n] u8 amps actually sent (schedule of env.py:368-378) or NULL
 *   final_obs: with AUTORESET, receives the terminal observation of envs that finished, or NULL
 *   status : [B] u8 or NULL;  stats: [16] f64 accumulators or NULL (see sg_stats_*)
 *   work   : [SG_EV_WORK_INTS(B)] i32 zero-initialised device scratch (tile scheduler + work list of the exact
 *            projection); required when m->project != 0, may be NULL otherwise (tiles are then assigned by a
 *            grid stride).  It is re-armed by every call.
 * With m->project != 0 the box-clipped action is tested against the network constraints per env (it
 * is then the exact projection); the few envs that violate them are finished by helper CTAs of the
 * same launch (exact single-class solve, else interior point) and report SG_EV_STATUS_PROJECTED.
 * Diagnostic environment variables (results never depend on them): SG_EV_MAX_CTAS=<n> caps the tile CTAs of the
 * launch; SG_EV_TILING="<full tiles>,<envs per tile after them>" overrides how the batch is cut into warp tiles
 * (default: 16-env tiles, the envs of a thin last round in 8-env tiles).
 */
int sg_ev_step(const sg_ev_model* m, const sg_ev_state* s, int64_t num_envs,
               const void* action, int32_t flags, int32_t discrete_bins,
               float* obs, double* reward, uint8_t* done, uint8_t* pilots, float* final_obs,
               uint8_t* status, double* stats, int32_t* work, void* stream);

/* Host-buffer variant: H2D(action) -> step -> D2H(obs, reward, done) -> sync. */
int sg_ev_step_host(const sg_ev_model* m, const sg_ev_state* s, int64_t num_envs,
                    const void* h_action, int32_t flags, int32_t discrete_bins,
                    float* h_obs, double* h_reward, uint8_t* h_done,
                    void* d_action, float* d_obs, double* d_reward, uint8_t* d_done,
                    uint8_t* d_status, double* d_stats, int32_t* d_work, void* stream);

/* Same without the final synchronisation (SB3 VecEnv.step_async): the host buffers must be pinned and
 * are valid after the caller synchronises `stream`.  Several env batches stepped this way on different
 * streams overlap their host->device and device->host copies. */
int sg_ev_step_host_async(const sg_ev_model* m, const sg_ev_state* s, int64_t num_envs,
                          const void* h_action, int32_t flags, int32_t discrete_bins,
                          float* h_obs, double* h_reward, uint8_t* h_done,
                          void* d_action, float* d_obs, double* d_reward, uint8_t* d_done,
                          uint8_t* d_status, double* d_stats, int32_t* d_work, void* stream);

/* ---- compact host path -------------------------------------------------------------------------------------
 * A caller that keeps its observations in host memory (SB3 VecEnv, examples/evcharging/train_stable_baselines.py:264-275)
 * does not need the dense [B, 2n+k+2] row to cross PCIe: ~6 of 54 stations hold an active session and the MOER
 * row is a table look-up.  After a step the device packs one 120-byte record per env, the records are copied to
 * the host and sg_ev_expand_obs rebuilds the observation there, bit-identical to the dense copy. */
#define SG_EV_COMPACT_ENTRIES      16   /* active sessions held in a record                                       */
#define SG_EV_COMPACT_OVERFLOW     1    /* flags bit: more active sessions than entries -> rows in the overflow area */
#define SG_EV_COMPACT_HEADER_BYTES 16   /* the compact buffer starts with { i32 n_overflow, 12 bytes padding }     */
#define SG_EV_COMPACT_BYTES(B)     (SG_EV_COMPACT_HEADER_BYTES + (size_t)(B) * 120)
typedef struct sg_ev_compact_obs {
    uint64_t active;       /* bit i: station i has an active session (non-zero demand / est_departure entries)  */
    double   reward;
    int16_t  t;            /* env clock after the step (0..288): timestep = t / 288, MOER row t                  */
    int16_t  moer_day;     /* MOER table index the observation reads                                             */
    uint8_t  done;
    uint8_t  flags;        /* SG_EV_COMPACT_OVERFLOW                                                             */
    uint8_t  n_active;
    uint8_t  status;       /* SG_EV_STATUS_* of the step                                                         */
    float    demand[SG_EV_COMPACT_ENTRIES];   /* active stations in ascending station order                       */
    int16_t  est[SG_EV_COMPACT_ENTRIES];      /* est_departure - t (integer-valued in the observation)            */
} sg_ev_compact_obs;

/* Packs the observation / reward / done / status of every env (as written by sg_ev_step or sg_ev_reset) into
 * `compact` ([device] SG_EV_COMPACT_BYTES(B), 16-byte aligned).  Envs with more than SG_EV_COMPACT_ENTRIES active
 * sessions set SG_EV_COMPACT_OVERFLOW and append their index to ovf_idx and their demand | est_departure rows
 * ([2n] f32) to ovf_rows (at most ovf_cap of them; the counter in the buffer header keeps counting). */
int sg_ev_pack_obs(int64_t num_envs, int32_t n_stations, int32_t obs_dim, const float* obs, const double* reward,
                   const uint8_t* done, const uint8_t* status /* or NULL */, const void* header /* sg_ev_state.header */,
                   void* compact, int32_t* ovf_idx /* [device] [ovf_cap] */, float* ovf_rows /* [device] [ovf_cap, 2n] */,
                   int32_t ovf_cap, void* stream);

/* H2D(action) -> step -> pack -> D2H(compact records), no synchronisation: h_action / h_compact must be pinned;
 * after the caller synchronises `stream`, h_compact holds SG_EV_COMPACT_BYTES(B) bytes for sg_ev_expand_obs. */
int sg_ev_step_host_compact_async(const sg_ev_model* m, const sg_ev_state* s, int64_t num_envs,
                                  const void* h_action, int32_t flags, int32_t discrete_bins, void* h_compact,
                                  void* d_action, float* d_obs, double* d_reward, uint8_t* d_done, uint8_t* d_status,
                                  double* d_stats, int32_t* d_work, void* d_compact, int32_t* d_ovf_idx,
                                  float* d_ovf_rows, int32_t ovf_cap,
                                  float* d_final_obs /* [device] [B, 2n+k+2] or NULL: terminal observations (AUTORESET) */,
                                  void* stream);

/* Host buffers of a stream-ordered expansion (sg_ev_step_host_compact_stream): everything sg_ev_expand_obs takes. */
typedef struct sg_ev_host_expand {
    const float* h_moer;        /* host copy of sg_ev_model.moer                                               */
    int32_t  n_moer_days;
    int32_t  incremental;       /* see sg_ev_expand_obs                                                         */
    uint64_t* h_prev_active;    /* [B] or NULL                                                                  */
    int32_t* h_ovf_idx;         /* [ovf_cap] pinned                                                             */
    float*   h_ovf_rows;        /* [ovf_cap, 2n] pinned                                                         */
    float*   h_obs;             /* [B, 2n+k+2]                                                                  */
    double*  h_reward;          /* [B]                                                                          */
    uint8_t* h_done;            /* [B]                                                                          */
    uint8_t* h_status;          /* [B] or NULL                                                                  */
    int32_t  n_threads;
    int32_t  reserved;
} sg_ev_host_expand;

/* sg_ev_step_host_compact_async + the overflow area copied to x->h_ovf_idx / x->h_ovf_rows + sg_ev_expand_obs enqueued
 * on `stream` as a host function (cudaLaunchHostFunc): the expansion runs on a driver thread as soon as the records
 * have arrived, overlapped with the caller and with other streams' work.  After the caller synchronises `stream` the
 * host buffers of `x` hold the step's results, unless the overflow counter (first int32 of h_compact) exceeds ovf_cap:
 * then the rows of the overflowing envs are incomplete and the caller copies the dense observation instead.  Overflow
 * indices are relative to the first compact env (dense_envs).  The
 * struct `x` is copied; the buffers it points to must stay valid until the stream has been synchronised. */
int sg_ev_step_host_compact_stream(const sg_ev_model* m, const sg_ev_state* s, int64_t num_envs,
                                   const void* h_action, int32_t flags, int32_t discrete_bins, void* h_compact,
                                   void* d_action, float* d_obs, double* d_reward, uint8_t* d_done, uint8_t* d_status,
                                   double* d_stats, int32_t* d_work, void* d_compact, int32_t* d_ovf_idx,
                                   float* d_ovf_rows, int32_t ovf_cap, float* d_final_obs,
                                   const sg_ev_host_expand* x,
                                   int64_t dense_envs /* the first dense_envs envs are copied as dense rows (pinned
                                                         x->h_obs / h_reward / h_done), the rest as records: balances the
                                                         PCIe link against the host's store bandwidth; 0 = all compact */,
                                   void* stream);

/* HOST function (no CUDA call): rebuilds obs [B, 2n+k+2] f32, reward [B] f64, done [B] u8 (and status [B] u8 unless
 * NULL) from the compact records with `n_threads` threads of a persistent pool.  `h_moer` is the host copy of
 * sg_ev_model.moer.  incremental != 0: `h_obs` still holds the previous expansion of the same envs and
 * `h_prev_active` ([B] u64, zero-initialised together with h_obs) its activity masks — only entries that changed
 * are written; incremental == 0 rewrites every row (h_prev_active may be NULL).  h_ovf_idx / h_ovf_rows: host
 * copies of the first n_ovf overflow entries (n_ovf = min(counter, ovf_cap); the caller falls back to a dense
 * copy when the counter exceeds ovf_cap). */
int sg_ev_expand_obs(const void* h_compact, int64_t num_envs, int32_t n_stations, int32_t moer_steps,
                     const float* h_moer, int32_t n_moer_days, uint64_t* h_prev_active, int32_t incremental,
                     const int32_t* h_ovf_idx, const float* h_ovf_rows, int32_t n_ovf,
                     float* h_obs, double* h_reward, uint8_t* h_done, uint8_t* h_status, int32_t n_threads);

/* Dense views of the sparse state: delivered [B,n] f64 (kWh delivered to the session plugged in at
 * each station, 0 where empty) and sess_idx [B,n] u8 (its trace slot, 255 where empty); either may be NULL. */
int sg_ev_export_state(const sg_ev_model* m, const sg_ev_state* s, int64_t num_envs,
                       double* delivered, uint8_t* sess_idx, void* stream);

/* The projection alone (env.py:200-221): x = argmin ||x-a||_2 s.t. 0<=x<=min(1,d/(c*32)),
 * 32*|A~ x| <= magnitudes.  a: [B,n] f32, demands: [B,n] f32, x out: [B,n] f64. */
int sg_ev_project(const sg_ev_model* m, int64_t num_envs, const float* action,
                  const float* demands, double* x_out, uint8_t* status, int32_t* iters,
                  void* stream);

/* MultiAgentEVChargingEnv observation fan-out as documented (sustaingym/envs/evcharging/multiagent_env.py:102-148):
 * out[b, i, :] = cur_obs[b, :] with the demand / est. departure entries of the OTHER stations (e < 2n, e != i, n+i)
 * taken from past_obs.
 * cur_obs/past_obs: [B, obs_dim]; out: [B, n, obs_dim]. */
int sg_ev_agent_obs(const float* cur_obs, const float* past_obs, int64_t num_envs, int32_t n_stations,
                    int32_t obs_dim, float* out, void* stream);

/* GMM day-trace sampler: GMMsTraceGenerator._sample / ._create_events
 * (sustaingym/envs/evcharging/event_generation.py:416-515) + the per-reset event packing (:149-207)
 * for a whole pool of days, written straight into the device layouts of sg_ev_model.sessions /
 * .arrivals.  The mixture is the reference's pickled scikit-learn model flattened by
 * tools/export_ev_gmms.py. */
typedef struct sg_ev_gmm {
    int32_t n_components;        /* 30 for the packaged models; <= 64                                     */
    int32_t n_counts;            /* length of `counts`                                                     */
    int32_t n_stations;          /* 54 (Caltech) / 52 (JPL)                                                */
    int32_t reserved;
    const double*  cum_weights;  /* [device] [nc] cumulative mixture weights                               */
    const double*  means;        /* [device] [nc, 4] arrival, departure, est. departure (fractions of a
                                    day), requested energy / 100 kWh                                       */
    const double*  chol;         /* [device] [nc, 4, 4] lower Cholesky factors of the covariances          */
    const int32_t* counts;       /* [device] [n_counts] sessions per day over the period (:479-480)        */
    const double*  station_p;    /* [device] [n] station usage / total usage (:493)                        */
} sg_ev_gmm;

/* Samples days first_day .. first_day + n_days - 1 of the stream `seed` (day d is a pure function of
 * (seed, d): Philox4x32-10, counter = (block, d)).  Outputs, all [device]:
 *   sessions [n_days, max_sessions, n] sg_ev_session; arrivals [n_days, max_day_sessions + 1] u32;
 *   dropped  [n_days] i32 or NULL: sessions that found no free station (:500-501) or no table slot;
 *   max_kwh  [n_days] f64 or NULL: sum of min(requested kWh, (departure - arrival) * kwh_per_step),
 *            the quantity behind EVChargingEnv._calculate_max_profit (env.py:422-429). */
int sg_ev_sample_gmm_days(const sg_ev_gmm* g, int64_t n_days, int64_t first_day, uint64_t seed,
                          double requested_energy_cap, double kwh_per_step, int32_t max_sessions,
                          int32_t max_day_sessions, void* sessions, uint32_t* arrivals, int32_t* dropped,
                          double* max_kwh, void* stream);

/* ------------------------------------------------------------------------------------------
 * CogenEnv  —  replaces sustaingym/envs/cogen/env.py:276-353 (_compute_reward, incl. the ONNX
 *   Runtime session.run at :305), :232-274 (_dyn_constraint_volation), :145-184 (_get_obs /
 *   _forecast_from_time), :355-387 (step), :186-230 (reset) and
 *   sustaingym/envs/cogen/multiagent_env.py:97-101 (per-agent reward split).
 * The plant model is the reference's model.onnx repacked by tests/golden/gen_cogen_pack.py:
 * 23 sub-models  (x-mu)*sinv -> Dense(d,H1) -> act -> Dense(H1,H2) -> act -> Dense(H2,1).
 * ---------------------------------------------------------------------------------------- */
#define SG_COGEN_ACTION_DIM 15   /* model inputs 3..17 (model.json order): GTi_{PAC,EVC}_FFU, GTi_PWR,
                                    HR1..3_HPIP_M_PROC, ST_PWR, IPPROC_M, CT_NrBays                */
#define SG_COGEN_N_SUB      23
#define SG_COGEN_STEPS      96

typedef struct sg_cogen_model {
    int32_t n_sub;                       /* 23                                                        */
    int32_t forecast_horizon;            /* k (cogen/env.py:80)                                       */
    int32_t n_days;                      /* days in the ambient table                                 */
    int32_t reserved;
    double  ramp_penalty;                /* cogen/env.py:77-79                                        */
    double  supply_imbalance_penalty;
    double  constraint_violation_penalty;
    const float*   blob;     /* [device] packed f32 weights (W1,b1,W2,b2,W3 per sub-model, 16 B aligned) */
    const int64_t* offsets;  /* [device] [23, 5] element offsets of W1,b1,W2,b2,W3 in blob              */
    const int32_t* meta;     /* [device] [23, 6] d, H1, H2, act (0 sigmoid / 1 relu), y index, gate     */
    const int32_t* feats;    /* [device] [23, 16] indices into the extended feature vector, -1 padded   */
    const float*   mu;       /* [device] [23, 16]                                                       */
    const float*   sinv;     /* [device] [23, 16]                                                       */
    const float*   b3;       /* [device] [23]                                                           */
    const double*  ambient;  /* [device] [n_days, 96, 7]: TAMB, PAMB, RHAMB, Target_Power, Target_Steam,
                                Energy_Price, Gas_Price (load_ambients.construct_df columns)            */
    const float*   tc_blob;    /* [device] pre-split (3xTF32) layer-2 weights for the tensor-core kernel, filled
                                  by sg_cogen_prepare_tc; NULL selects the fp32 FFMA kernel                    */
    const int64_t* tc_offsets; /* [device] [23] float offset of each sub-model in tc_blob:
                                  cumulative (H1/16) * (272 + 32*H2) floats                                   */
} sg_cogen_model;

/* Fills tc_blob (m->tc_offsets must be set; sized by its last offset + last sub-model) from the fp32
 * weights: per 16-unit chunk W1 | b1 | W2 hi | W2 lo in the tcgen05 K-major operand layout. */
int sg_cogen_prepare_tc(const sg_cogen_model* m, float* tc_blob, void* stream);

/* Device scratch of the plant model, in floats, 16-byte aligned: the 23 raw sub-model outputs per env,
 * plus (tensor-core path) the normalised, zero-padded inputs of every sub-model [23,B,16]. */
#define SG_COGEN_SCRATCH_FLOATS(B) ((int64_t)(B) * (23 + 23 * 16) + 4)

/* The plant model alone: X [B,18] f32 -> Y [B,29] f32 (model.json order).
 * scratch: SG_COGEN_SCRATCH_FLOATS(B) f32. */
int sg_cogen_plant(const sg_cogen_model* m, int64_t num_envs, const float* X, float* Y, float* scratch,
                   void* stream);

/* reset(): day_in [B] picks the day (seed -> day map is host logic, cogen/env.py:213-216);
 * init_action [B,15] is the randomly drawn initial Prev_Action (cogen/env.py:222-223, host RNG). */
int sg_cogen_reset(const sg_cogen_model* m, int64_t num_envs, const int32_t* day_in, const float* init_action,
                   float* prev_action /* [B,15] out */, int32_t* day_t /* [B,2] out */,
                   float* obs /* [B, 7(k+1)+33] out */, void* stream);

/* One step() of B envs. action/prev_action: [B,15] f32; day_t: [B,2] {day, t}; noise: [B,k,7] f64
 * forecast noise already scaled by forecast_noise_std, or NULL; obs: flattened Dict (sorted keys);
 * reward [B] f64; agent_reward [B,4] f64 (GT1, GT2, GT3, ST); done [B]; Y_out [B,29] or NULL. */
int sg_cogen_step(const sg_cogen_model* m, int64_t num_envs, const float* action, float* prev_action,
                  int32_t* day_t, const double* noise, float* obs, double* reward, double* agent_reward,
                  uint8_t* done, float* Y_out, float* scratch /* SG_COGEN_SCRATCH_FLOATS(B) */, int32_t flags, double* stats,
                  void* stream);

/* Unit test of the tensor-core path of the plant model: D[128,N] = A[128,K] * Bm[N,K]^T through 3xTF32
 * tcgen05.mma (operands split in shared memory, fp32 accumulator in tensor memory). N in {128, 256},
 * K a multiple of 8, <= 64.  All pointers [device], row-major. */
int sg_tc_gemm_selftest(const float* A, const float* Bm, float* D, int32_t N, int32_t K, void* stream);

/* ------------------------------------------------------------------------------------------
 * ElectricityMarketEnv  —  SPEC-DEFINED: the reference snapshot has no source for it (only
 *   docs/electricitymarketenv.md:9-27).  Spec: sustaingym_b200/envs/electricitymarket/SPEC.md.
 * ---------------------------------------------------------------------------------------- */
typedef struct sg_market_model {
    int32_t n_gen;                 /* generators (<= 64), costs ascending                            */
    int32_t k;                     /* bid / forecast horizon (1..4)                                   */
    int32_t n_days;                /* rows of the load / MOER day tables                              */
    int32_t use_intermediate_rewards;
    double  e_max, e_init, p_max, eta_c, eta_d, price_cap, p_co2, terminal_price;
    const double* gen_cost;        /* [device] [n_gen] $/MWh                                          */
    const double* gen_cap;         /* [device] [n_gen] MW                                             */
    const double* load;            /* [device] [n_days, 288 + k] MW                                   */
    const double* moer;            /* [device] [n_days, 288 + k] kg CO2 / MWh                         */
} sg_market_model;

int sg_market_reset(const sg_market_model* m, int64_t num_envs, const int32_t* day_in,
                    double* state /* [B,4]: e, x_prev, p_prev, deferred reward */, int32_t* day_t /* [B,2] */,
                    float* prev_action /* [B,2k] */, float* obs /* [B, 6+4k] */, void* stream);
int sg_market_step(const sg_market_model* m, int64_t num_envs, const float* action /* [B,2k] */, double* state,
                   int32_t* day_t, float* prev_action, float* obs, double* reward, uint8_t* done,
                   double* info /* [B,3] price, dispatch, energy or NULL */, double* stats, void* stream);

/* ------------------------------------------------------------------------------------------
 * Aggregate reward statistics (the one collective of the multi-GPU path):
 * kernels accumulate into a 16-double vector; the host all-reduces it (NCCL SUM / MIN / MAX).
 *   [0] count  [1] sum r  [2] sum r^2  [3] min r  [4] max r  [5] n_done  [6] n_projected
 *   [7] n_maxiter  [8..10] sum of reward components  [11..15] reserved
 * ---------------------------------------------------------------------------------------- */
#define SG_STATS_LEN 16
int sg_stats_reset(double* stats /* [device] [16] */, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SUSTAINGYM_B200_H */

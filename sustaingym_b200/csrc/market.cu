// ElectricityMarketEnv batched step()/reset() for sm_100a — SPEC-DEFINED (no reference source exists;
// see sustaingym_b200/envs/electricitymarket/SPEC.md and docs/electricitymarketenv.md:9-27 of the
// reference).  One thread per env: single-bus merit-order clearing against a fleet of <= 64
// generators staged in shared memory (costs ascending, cumulative capacities), battery SoC update,
// revenue + carbon reward, flat observation.  Bound: HBM (state 64 B r/w + action + obs per env-step).
#include "common.cuh"

namespace sg {
namespace {

constexpr int kT = 288;
constexpr double kTau = 5.0 / 60.0;
constexpr int kMaxGen = 64;

// per-env state row (8 doubles = 64 B): e, x_prev, p_prev, acc, t, day, a_c_prev.., kept as doubles
struct MarketShared { double cost[kMaxGen], cap[kMaxGen], cum[kMaxGen]; double stat[11 * 4]; };

__device__ __forceinline__ void write_market_obs(const sg_market_model& m, const int day, const int t, const double e,
                                                 const float* prev_a, const double x_prev, const double p_prev,
                                                 float* o) {
    const int k = m.k, W = kT + k;
    const double* lrow = m.load + (long long)day * W;
    const double* mrow = m.moer + (long long)day * W;
    int pos = 0;
    o[pos++] = (float)t; o[pos++] = (float)e;
    for (int i = 0; i < 2 * k; ++i) o[pos++] = prev_a[i];
    o[pos++] = (float)x_prev; o[pos++] = (float)p_prev;
    o[pos++] = t > 0 ? (float)lrow[t - 1] : 0.f;
    for (int i = 0; i < k; ++i) o[pos++] = (float)lrow[min(t + i, W - 1)];
    o[pos++] = t > 0 ? (float)mrow[t - 1] : 0.f;
    for (int i = 0; i < k; ++i) o[pos++] = (float)mrow[min(t + i, W - 1)];
}

__global__ void __launch_bounds__(128)
market_step_kernel(const sg_market_model m, const long long B, const float* __restrict__ action,
                   double* __restrict__ state /* [B,4]: e, x_prev, p_prev, acc */, int32_t* __restrict__ day_t,
                   float* __restrict__ prev_a /* [B,2k] */, float* __restrict__ obs, double* __restrict__ reward,
                   uint8_t* __restrict__ done, double* __restrict__ info /* [B,3] price, dispatch, energy or NULL */,
                   double* __restrict__ stats) {
    __shared__ MarketShared sh;
    // A step is one wave of CTAs: its duration is the chain of dependent memory round trips, so every load is requested
    // as early as its address is known — the env's clock first (the load / MOER rows depend on it), then the fleet
    // tables, the env's state and bids, then the load / MOER values — and the first barrier comes after all of them.
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int k = m.k, W = kT + k, G = m.n_gen, obs_dim = 6 + 4 * k;
    int day = 0, t = kT;
    if (b < B) { day = day_t[b * 2]; t = day_t[b * 2 + 1]; }
    for (int i = threadIdx.x; i < kMaxGen; i += blockDim.x) {
        sh.cost[i] = i < m.n_gen ? m.gen_cost[i] : 1e300;
        sh.cap[i] = i < m.n_gen ? m.gen_cap[i] : 0.0;
    }
    double e = 0.0, racc = 0.0, load = 0.0, moer = 0.0;
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const bool stepping = b < B && t < kT;
    if (stepping) {
        e = state[b * 4 + 0]; racc = state[b * 4 + 3];
        for (int i = 0; i < 2 * k; ++i) a[i] = fmaxf(action[b * 2 * k + i], 0.f);
        load = m.load[(long long)day * W + t]; moer = m.moer[(long long)day * W + t];
    }
    __syncthreads();
    // cumulative capacities: thread i adds cap[0..i] in index order (the rounding of a sequential running sum, computed
    // by 64 threads side by side instead of one thread walking the array between two barriers)
    if (threadIdx.x < kMaxGen) {
        double c = 0;
        for (int i = 0; i <= (int)threadIdx.x; ++i) c += sh.cap[i];
        sh.cum[threadIdx.x] = c;
    }
    __syncthreads();
    StatsAcc acc;
    if (b < B) {
        if (t >= kT) { reward[b] = 0.0; done[b] = 1; }
        else {
            const double a_c = (double)a[0], a_d = (double)a[k];
            const double q_d = fmax(fmin(m.p_max, e * m.eta_d / kTau), 0.0);
            const double q_c = fmax(fmin(m.p_max, (m.e_max - e) / (m.eta_c * kTau)), 0.0);
            double cheap_d = 0.0, cheap_c = 0.0;
            for (int g = 0; g < G; ++g) {
                cheap_d += sh.cost[g] <= a_d ? sh.cap[g] : 0.0;
                cheap_c += sh.cost[g] <= a_c ? sh.cap[g] : 0.0;
            }
            const double x_d = fmin(fmax(load - cheap_d, 0.0), q_d);
            const double x_c = x_d > 0.0 ? 0.0 : fmin(fmax(cheap_c - load, 0.0), q_c);
            const double net = load - x_d + x_c;
            int idx = 0;
            for (int g = 0; g < G; ++g) idx += sh.cum[g] < net - 1e-12 ? 1 : 0;
            double gen_price = idx < G ? sh.cost[idx] : m.price_cap;
            if (net <= 0.0) gen_price = 0.0;
            const double price = (x_d > 0.0 && x_d < q_d) ? a_d : ((x_c > 0.0 && x_c < q_c) ? a_c : gen_price);
            const double x = (x_d - x_c) * kTau;
            e = fmin(fmax(e - x_d * kTau / m.eta_d + x_c * kTau * m.eta_c, 0.0), m.e_max);
            double r = price * x + m.p_co2 * moer * x;
            const int tn = t + 1;
            const bool term = tn >= kT;
            if (term) r -= m.terminal_price * fmax(0.0, m.e_init - e);
            double rew = r;
            if (!m.use_intermediate_rewards) { racc += r; rew = term ? racc : 0.0; }
            state[b * 4 + 0] = e; state[b * 4 + 1] = x; state[b * 4 + 2] = price; state[b * 4 + 3] = racc;
            for (int i = 0; i < 2 * k; ++i) prev_a[b * 2 * k + i] = a[i];
            day_t[b * 2 + 1] = tn;
            write_market_obs(m, day, tn, e, a, x, price, obs + b * obs_dim);
            reward[b] = rew; done[b] = term ? 1 : 0;
            if (info) { info[b * 3] = price; info[b * 3 + 1] = x; info[b * 3 + 2] = e; }
            acc.add(rew, term, price * x, m.p_co2 * moer * x, 0.0);
        }
    }
    if (stats != nullptr) {
        __syncthreads();
        stats_flush_block(acc, stats, sh.stat);
    }
}

__global__ void market_reset_kernel(const sg_market_model m, const long long B, const int32_t* __restrict__ day_in,
                                    double* __restrict__ state, int32_t* __restrict__ day_t,
                                    float* __restrict__ prev_a, float* __restrict__ obs) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int k = m.k;
    const int day = min(max(day_in[b], 0), m.n_days - 1);
    state[b * 4 + 0] = m.e_init; state[b * 4 + 1] = 0.0; state[b * 4 + 2] = 0.0; state[b * 4 + 3] = 0.0;
    day_t[b * 2] = day; day_t[b * 2 + 1] = 0;
    float a[8];
    for (int i = 0; i < 2 * k; ++i) { a[i] = 0.f; prev_a[b * 2 * k + i] = 0.f; }
    write_market_obs(m, day, 0, m.e_init, a, 0.0, 0.0, obs + b * (6 + 4 * k));
}

int check_market(const sg_market_model* m) {
    SG_CHECK_ARG(m != nullptr, "market: model is NULL");
    SG_CHECK_ARG(m->n_gen >= 1 && m->n_gen <= kMaxGen, "market: n_gen=%d outside [1,64]", m->n_gen);
    SG_CHECK_ARG(m->k >= 1 && m->k <= 4, "market: k=%d outside [1,4]", m->k);
    SG_CHECK_ARG(m->gen_cost && m->gen_cap && m->load && m->moer && m->n_days >= 1, "market: NULL/empty table");
    return 0;
}

}  // namespace
}  // namespace sg

extern "C" {

int sg_market_reset(const sg_market_model* m, int64_t num_envs, const int32_t* day_in, double* state,
                    int32_t* day_t, float* prev_action, float* obs, void* stream) {
    if (int rc = sg::check_market(m)) return rc;
    SG_CHECK_ARG(day_in && state && day_t && prev_action && obs, "market_reset: NULL argument");
    if (num_envs <= 0) return 0;
    sg::market_reset_kernel<<<(unsigned)((num_envs + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        *m, num_envs, day_in, state, day_t, prev_action, obs);
    return sg::launch_status("market_reset_kernel");
}

int sg_market_step(const sg_market_model* m, int64_t num_envs, const float* action, double* state, int32_t* day_t,
                   float* prev_action, float* obs, double* reward, uint8_t* done, double* info, double* stats,
                   void* stream) {
    if (int rc = sg::check_market(m)) return rc;
    SG_CHECK_ARG(action && state && day_t && prev_action && obs && reward && done, "market_step: NULL argument");
    if (num_envs <= 0) return 0;
    sg::market_step_kernel<<<(unsigned)((num_envs + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        *m, num_envs, action, state, day_t, prev_action, obs, reward, done, info, stats);
    return sg::launch_status("market_step_kernel");
}

}  // extern "C"

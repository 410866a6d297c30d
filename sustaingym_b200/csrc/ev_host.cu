// Host side of the EVChargingEnv host-buffer path: compact observation records, packed on the device after a
// step (ev_pack_obs_kernel), cross PCIe instead of the dense [B, 2n+k+2] f32 observation (sparse: ~6 of 54
// stations hold an active session; the MOER row is a table look-up the host can do itself), and are expanded
// into the caller's observation buffer by a persistent pool of host threads (sg_ev_expand_obs), bit-identical
// to a device->host copy of the dense observation (tests/test_ev_gpu.py::test_compact_host_path_*).
//
// Replaces, for callers that keep their observations in host memory (StableBaselines3 VecEnv, the reference's
// training scripts, examples/evcharging/train_stable_baselines.py:264-275), the device->host copy of
//   EVChargingEnv._get_observation   sustaingym/envs/evcharging/env.py:381-394
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

#if defined(__x86_64__)
#include <emmintrin.h>
#define SG_HAVE_STREAM_STORES 1
#else
#define SG_HAVE_STREAM_STORES 0
#endif

namespace sg {
namespace {

static_assert(sizeof(sg_ev_compact_obs) == 120, "sg_ev_compact_obs must be 120 bytes");

struct EnvHeaderView {          // first 16 bytes of the 64-byte env record (csrc/ev.cu EnvHeader)
    int t, day, moer_day, nidx;
};

// One warp per env: the dense observation row the step kernel just wrote (L2-resident) -> compact record.
__global__ void __launch_bounds__(256)
ev_pack_obs_kernel(const long long B, const int n, const int obs_dim, const float* __restrict__ obs,
                   const double* __restrict__ reward, const uint8_t* __restrict__ done,
                   const uint8_t* __restrict__ status, const unsigned char* __restrict__ header,
                   sg_ev_compact_obs* __restrict__ rec, int32_t* __restrict__ ovf_count,
                   int32_t* __restrict__ ovf_idx, float* __restrict__ ovf_rows, const int ovf_cap) {
    const int lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const float* row = obs + b * obs_dim;
    const int i0 = lane, i1 = lane + 32;
    const float d0 = i0 < n ? row[i0] : 0.f, d1 = i1 < n ? row[i1] : 0.f;
    const float e0 = i0 < n ? row[n + i0] : 0.f, e1 = i1 < n ? row[n + i1] : 0.f;
    // an active session has a remaining demand > 1e-3 kWh (env.py:388 through Interface.active_sessions); every
    // other station reads 0 in both fields
    const unsigned m0 = __ballot_sync(kFull, d0 != 0.f || e0 != 0.f), m1 = __ballot_sync(kFull, d1 != 0.f || e1 != 0.f);
    const unsigned lt = (1u << lane) - 1u;
    const int c0 = __popc(m0), cnt = c0 + __popc(m1);
    sg_ev_compact_obs* r = rec + b;
    const bool over = cnt > SG_EV_COMPACT_ENTRIES;
    if ((m0 >> lane) & 1u) {
        const int k = __popc(m0 & lt);
        if (k < SG_EV_COMPACT_ENTRIES) { r->demand[k] = d0; r->est[k] = (int16_t)__float2int_rn(e0); }
    }
    if ((m1 >> lane) & 1u) {
        const int k = c0 + __popc(m1 & lt);
        if (k < SG_EV_COMPACT_ENTRIES) { r->demand[k] = d1; r->est[k] = (int16_t)__float2int_rn(e1); }
    }
    int slot = -1;
    if (over) {                                    // rare: the dense demand / departure rows go to the overflow area
        if (lane == 0) slot = atomicAdd(ovf_count, 1);
        slot = __shfl_sync(kFull, slot, 0);
        if (slot < ovf_cap) {
            if (lane == 0) ovf_idx[slot] = (int32_t)b;
            float* o = ovf_rows + (long long)slot * 2 * n;
            for (int i = lane; i < 2 * n; i += 32) o[i] = row[i];
        }
    }
    if (lane == 0) {
        const EnvHeaderView h = *reinterpret_cast<const EnvHeaderView*>(header + b * 64);
        r->active = ((unsigned long long)m1 << 32) | m0;
        r->reward = reward[b];
        r->t = (int16_t)h.t;
        r->moer_day = (int16_t)h.moer_day;
        r->done = done[b];
        r->flags = over ? SG_EV_COMPACT_OVERFLOW : 0;
        r->n_active = (uint8_t)cnt;
        r->status = status != nullptr ? status[b] : 0;
    }
}

// ---- persistent host thread pool --------------------------------------------------------------------------------
class Pool {
public:
    static Pool& get() { static Pool p; return p; }

    // runs fn(part, nparts) for part in [0, nparts) on nparts threads (the caller is one of them)
    void run(int nparts, const std::function<void(int, int)>& fn) {
        if (nparts <= 1) { fn(0, 1); return; }
        std::unique_lock<std::mutex> call(call_mu_);            // one parallel region at a time
        grow(nparts - 1);
        {
            std::lock_guard<std::mutex> g(mu_);
            fn_ = &fn; nparts_ = nparts; pending_ = nparts - 1; ++gen_;
        }
        cv_.notify_all();
        fn(0, nparts);
        std::unique_lock<std::mutex> g(mu_);
        done_cv_.wait(g, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    Pool() = default;
    ~Pool() {
        { std::lock_guard<std::mutex> g(mu_); stop_ = true; ++gen_; }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    void grow(int n) {
        while ((int)threads_.size() < n) {
            const int id = (int)threads_.size() + 1;
            unsigned long long seen;
            { std::lock_guard<std::mutex> g(mu_); seen = gen_; }
            threads_.emplace_back([this, id, seen]() mutable { loop(id, seen); });
        }
    }
    void loop(int id, unsigned long long seen) {
        for (;;) {
            const std::function<void(int, int)>* fn;
            int nparts;
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                fn = fn_; nparts = nparts_;
            }
            if (fn != nullptr && id < nparts) {
                (*fn)(id, nparts);
                std::lock_guard<std::mutex> g(mu_);
                if (--pending_ == 0) done_cv_.notify_one();
            }
        }
    }
    std::mutex mu_, call_mu_;
    std::condition_variable cv_, done_cv_;
    std::vector<std::thread> threads_;
    const std::function<void(int, int)>* fn_ = nullptr;
    int nparts_ = 0, pending_ = 0;
    unsigned long long gen_ = 0;
    bool stop_ = false;
};

inline float timestep_frac_host(int t) { return (float)((double)t * (1.0 / (double)SG_EV_EPISODE_STEPS)); }

}  // namespace
}  // namespace sg

extern "C" {

int sg_ev_pack_obs(int64_t num_envs, int32_t n_stations, int32_t obs_dim, const float* obs, const double* reward,
                   const uint8_t* done, const uint8_t* status, const void* header, void* compact,
                   int32_t* ovf_idx, float* ovf_rows, int32_t ovf_cap, void* stream) {
    SG_CHECK_ARG(obs && reward && done && header && compact, "ev_pack_obs: NULL argument");
    SG_CHECK_ARG(n_stations >= 1 && n_stations <= SG_EV_MAX_STATIONS && obs_dim >= 2 * n_stations, "ev_pack_obs: bad dimensions");
    SG_CHECK_ARG(ovf_cap == 0 || (ovf_idx && ovf_rows), "ev_pack_obs: NULL overflow area");
    SG_CHECK_ARG((reinterpret_cast<uintptr_t>(compact) & 15u) == 0, "ev_pack_obs: compact buffer must be 16-byte aligned");
    if (num_envs <= 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned char* c = static_cast<unsigned char*>(compact);
    SG_CUDA(cudaMemsetAsync(c, 0, SG_EV_COMPACT_HEADER_BYTES, st));                 // overflow counter
    const unsigned grid = (unsigned)((num_envs + 7) / 8);
    sg::ev_pack_obs_kernel<<<grid, 256, 0, st>>>(num_envs, n_stations, obs_dim, obs, reward, done, status,
                                                 static_cast<const unsigned char*>(header),
                                                 reinterpret_cast<sg_ev_compact_obs*>(c + SG_EV_COMPACT_HEADER_BYTES),
                                                 reinterpret_cast<int32_t*>(c), ovf_idx, ovf_rows, ovf_cap);
    return sg::launch_status("ev_pack_obs_kernel");
}

int sg_ev_step_host_compact_async(const sg_ev_model* m, const sg_ev_state* s, int64_t num_envs, const void* h_action,
                                  int32_t flags, int32_t discrete_bins, void* h_compact, void* d_action, float* d_obs,
                                  double* d_reward, uint8_t* d_done, uint8_t* d_status, double* d_stats, int32_t* d_work,
                                  void* d_compact, int32_t* d_ovf_idx, float* d_ovf_rows, int32_t ovf_cap,
                                  float* d_final_obs, void* stream) {
    SG_CHECK_ARG(m && s && h_action && h_compact && d_action && d_obs && d_reward && d_done && d_compact,
                 "ev_step_host_compact: NULL argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t B = (size_t)num_envs, n = (size_t)m->n_stations;
    const int obs_dim = 2 * m->n_stations + m->moer_steps + 2;
    const size_t abytes = B * n * (((flags & 1) == SG_EV_ACTION_I64) ? 8 : 4);
    SG_CUDA(cudaMemcpyAsync(d_action, h_action, abytes, cudaMemcpyHostToDevice, st));
    if (int rc = sg_ev_step(m, s, num_envs, d_action, flags, discrete_bins, d_obs, d_reward, d_done, nullptr, d_final_obs,
                            d_status, d_stats, d_work, stream)) return rc;
    if (int rc = sg_ev_pack_obs(num_envs, m->n_stations, obs_dim, d_obs, d_reward, d_done, d_status, s->header, d_compact,
                                d_ovf_idx, d_ovf_rows, ovf_cap, stream)) return rc;
    SG_CUDA(cudaMemcpyAsync(h_compact, d_compact, SG_EV_COMPACT_HEADER_BYTES + B * sizeof(sg_ev_compact_obs),
                            cudaMemcpyDeviceToHost, st));
    return 0;
}

namespace {
struct ExpandJob {
    sg_ev_host_expand x;
    const void* h_compact;
    int64_t num_envs;
    int32_t n, k, ovf_cap;
};
void CUDART_CB expand_job(void* p) {
    ExpandJob* j = static_cast<ExpandJob*>(p);
    const int32_t n_ovf = *static_cast<const int32_t*>(j->h_compact);
    // (a counter beyond the overflow area: the caller falls back to the dense copy; expand what is there)
    sg_ev_expand_obs(j->h_compact, j->num_envs, j->n, j->k, j->x.h_moer, j->x.n_moer_days, j->x.h_prev_active,
                     j->x.incremental, j->x.h_ovf_idx, j->x.h_ovf_rows, n_ovf > j->ovf_cap ? 0 : n_ovf, j->x.h_obs,
                     j->x.h_reward, j->x.h_done, j->x.h_status, j->x.n_threads);
    delete j;
}
}  // namespace

int sg_ev_step_host_compact_stream(const sg_ev_model* m, const sg_ev_state* s, int64_t num_envs, const void* h_action,
                                   int32_t flags, int32_t discrete_bins, void* h_compact, void* d_action, float* d_obs,
                                   double* d_reward, uint8_t* d_done, uint8_t* d_status, double* d_stats, int32_t* d_work,
                                   void* d_compact, int32_t* d_ovf_idx, float* d_ovf_rows, int32_t ovf_cap,
                                   float* d_final_obs, const sg_ev_host_expand* x, int64_t dense_envs, void* stream) {
    SG_CHECK_ARG(m && s && h_action && h_compact && d_action && d_obs && d_reward && d_done && d_compact,
                 "ev_step_host_compact_stream: NULL argument");
    SG_CHECK_ARG(x && x->h_moer && x->h_obs && x->h_reward && x->h_done, "ev_step_host_compact_stream: NULL expansion buffer");
    SG_CHECK_ARG(ovf_cap == 0 || (x->h_ovf_idx && x->h_ovf_rows), "ev_step_host_compact_stream: NULL host overflow area");
    SG_CHECK_ARG(dense_envs >= 0 && dense_envs <= num_envs, "ev_step_host_compact_stream: dense_envs outside [0, num_envs]");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t B = (size_t)num_envs, n = (size_t)m->n_stations, dn = (size_t)dense_envs;
    const int obs_dim = 2 * m->n_stations + m->moer_steps + 2;
    const size_t abytes = B * n * (((flags & 1) == SG_EV_ACTION_I64) ? 8 : 4);
    SG_CUDA(cudaMemcpyAsync(d_action, h_action, abytes, cudaMemcpyHostToDevice, st));
    if (int rc = sg_ev_step(m, s, num_envs, d_action, flags, discrete_bins, d_obs, d_reward, d_done, nullptr, d_final_obs,
                            d_status, d_stats, d_work, stream)) return rc;
    // The first dense_envs envs travel as they are (the DMA engine writes their rows), the others as compact
    // records expanded by the host threads: the two run side by side, so the split balances the PCIe link against
    // the host's store bandwidth.
    const size_t rest = B - dn;
    if (rest > 0) {
        if (int rc = sg_ev_pack_obs((int64_t)rest, m->n_stations, obs_dim, d_obs + dn * obs_dim, d_reward + dn, d_done + dn,
                                    d_status ? d_status + dn : nullptr, static_cast<const unsigned char*>(s->header) + dn * 64,
                                    d_compact, d_ovf_idx, d_ovf_rows, ovf_cap, stream)) return rc;
        SG_CUDA(cudaMemcpyAsync(h_compact, d_compact, SG_EV_COMPACT_HEADER_BYTES + rest * sizeof(sg_ev_compact_obs),
                                cudaMemcpyDeviceToHost, st));
        if (ovf_cap > 0) {      // small (cap x 2n floats): copied every step so that the host function needs no CUDA call
            SG_CUDA(cudaMemcpyAsync(x->h_ovf_idx, d_ovf_idx, sizeof(int32_t) * ovf_cap, cudaMemcpyDeviceToHost, st));
            SG_CUDA(cudaMemcpyAsync(x->h_ovf_rows, d_ovf_rows, sizeof(float) * 2 * n * (size_t)ovf_cap, cudaMemcpyDeviceToHost, st));
        }
        ExpandJob* j = new ExpandJob{*x, h_compact, (int64_t)rest, m->n_stations, m->moer_steps, ovf_cap};
        j->x.h_obs += dn * obs_dim; j->x.h_reward += dn; j->x.h_done += dn;
        if (j->x.h_status) j->x.h_status += dn;
        if (j->x.h_prev_active) j->x.h_prev_active += dn;
        SG_CUDA(cudaLaunchHostFunc(st, expand_job, j));
    } else {
        SG_CUDA(cudaMemsetAsync(d_compact, 0, SG_EV_COMPACT_HEADER_BYTES, st));
        SG_CUDA(cudaMemcpyAsync(h_compact, d_compact, SG_EV_COMPACT_HEADER_BYTES, cudaMemcpyDeviceToHost, st));
    }
    if (dn > 0) {               // (queued behind the records so that the expansion starts while these rows still travel)
        SG_CUDA(cudaMemcpyAsync(x->h_obs, d_obs, dn * obs_dim * 4, cudaMemcpyDeviceToHost, st));
        SG_CUDA(cudaMemcpyAsync(x->h_reward, d_reward, dn * 8, cudaMemcpyDeviceToHost, st));
        SG_CUDA(cudaMemcpyAsync(x->h_done, d_done, dn, cudaMemcpyDeviceToHost, st));
        if (x->h_status && d_status) SG_CUDA(cudaMemcpyAsync(x->h_status, d_status, dn, cudaMemcpyDeviceToHost, st));
    }
    return 0;
}

int sg_ev_expand_obs(const void* h_compact, int64_t num_envs, int32_t n_stations, int32_t moer_steps,
                     const float* h_moer, int32_t n_moer_days, uint64_t* h_prev_active, int32_t incremental,
                     const int32_t* h_ovf_idx, const float* h_ovf_rows, int32_t n_ovf,
                     float* h_obs, double* h_reward, uint8_t* h_done, uint8_t* h_status, int32_t n_threads) {
    SG_CHECK_ARG(h_compact && h_moer && h_obs && h_reward && h_done, "ev_expand_obs: NULL argument");
    SG_CHECK_ARG(n_stations >= 1 && n_stations <= SG_EV_MAX_STATIONS && moer_steps >= 1 && moer_steps <= 36,
                 "ev_expand_obs: bad dimensions");
    SG_CHECK_ARG(!incremental || h_prev_active, "ev_expand_obs: incremental expansion needs the previous activity masks");
    SG_CHECK_ARG(n_ovf == 0 || (h_ovf_idx && h_ovf_rows), "ev_expand_obs: NULL overflow rows");
    if (num_envs <= 0) return 0;
    const int n = n_stations, k = moer_steps, obs_dim = 2 * n + k + 2;
    const sg_ev_compact_obs* rec = reinterpret_cast<const sg_ev_compact_obs*>(
        static_cast<const unsigned char*>(h_compact) + SG_EV_COMPACT_HEADER_BYTES);
    std::atomic<int> bad{0};
    // Full rewrite with streaming (non-temporal) stores: every row is composed in a small local buffer and leaves
    // the core as whole write-combined lines — no read-for-ownership of the destination, i.e. half the memory
    // traffic of cached stores (the 38 MB observation batch is far bigger than the caches and every one of a row's
    // cache lines changes each step anyway: MOER tail + scattered demand / departure entries).
    const bool stream_rows = SG_HAVE_STREAM_STORES && !incremental && ((obs_dim & 1) == 0) &&
                             (reinterpret_cast<uintptr_t>(h_obs) & 7u) == 0;
    auto body_stream = [&](int part, int nparts) {
#if SG_HAVE_STREAM_STORES
        const int64_t lo = num_envs * part / nparts, hi = num_envs * (part + 1) / nparts;
        alignas(64) float buf[2 * SG_EV_MAX_STATIONS + 40];
        for (int64_t b = lo; b < hi; ++b) {
            const sg_ev_compact_obs& r = rec[b];
            const uint64_t act = r.active;
            std::memset(buf, 0, sizeof(float) * 2 * n);
            if (!(r.flags & SG_EV_COMPACT_OVERFLOW)) {
                int j = 0;
                for (uint64_t mk = act; mk; mk &= mk - 1, ++j) {
                    const int i = __builtin_ctzll(mk);
                    buf[i] = r.demand[j]; buf[n + i] = (float)r.est[j];
                }
            }
            if (h_prev_active) h_prev_active[b] = act;
            const int md = r.moer_day, t = r.t;
            if (md < 0 || md >= n_moer_days || t < 0 || t > SG_EV_EPISODE_STEPS) { bad.store(1); continue; }
            const float* mrow = h_moer + ((int64_t)md * SG_EV_MOER_ROWS + t) * SG_EV_MOER_COLS;
            std::memcpy(buf + 2 * n, mrow + 1, sizeof(float) * k);
            buf[2 * n + k] = mrow[0];
            buf[2 * n + k + 1] = sg::timestep_frac_host(t);
            long long* dst = reinterpret_cast<long long*>(h_obs + b * obs_dim);
            const long long* src = reinterpret_cast<const long long*>(buf);
            for (int q = 0; q < obs_dim / 2; ++q) _mm_stream_si64(dst + q, src[q]);
            h_reward[b] = r.reward;
            h_done[b] = r.done;
            if (h_status) h_status[b] = r.status;
        }
        _mm_sfence();
#else
        (void)part; (void)nparts;
#endif
    };
    auto body = [&](int part, int nparts) {
        const int64_t lo = num_envs * part / nparts, hi = num_envs * (part + 1) / nparts;
        for (int64_t b = lo; b < hi; ++b) {
            const sg_ev_compact_obs& r = rec[b];
            float* row = h_obs + b * obs_dim;
            const uint64_t act = r.active;
            if (incremental) {
                for (uint64_t gone = h_prev_active[b] & ~act; gone; gone &= gone - 1) {
                    const int i = __builtin_ctzll(gone);
                    row[i] = 0.f; row[n + i] = 0.f;
                }
            } else {
                std::memset(row, 0, sizeof(float) * 2 * n);
            }
            if (!(r.flags & SG_EV_COMPACT_OVERFLOW)) {
                int j = 0;
                for (uint64_t mk = act; mk; mk &= mk - 1, ++j) {
                    const int i = __builtin_ctzll(mk);
                    row[i] = r.demand[j]; row[n + i] = (float)r.est[j];
                }
            }
            if (h_prev_active) h_prev_active[b] = act;
            const int md = r.moer_day, t = r.t;
            if (md < 0 || md >= n_moer_days || t < 0 || t > SG_EV_EPISODE_STEPS) { bad.store(1); continue; }
            const float* mrow = h_moer + ((int64_t)md * SG_EV_MOER_ROWS + t) * SG_EV_MOER_COLS;
            std::memcpy(row + 2 * n, mrow + 1, sizeof(float) * k);          // forecasted_moer
            row[2 * n + k] = mrow[0];                                         // prev_moer
            row[2 * n + k + 1] = sg::timestep_frac_host(t);                   // timestep
            h_reward[b] = r.reward;
            h_done[b] = r.done;
            if (h_status) h_status[b] = r.status;
        }
    };
    if (stream_rows) sg::Pool::get().run(n_threads < 1 ? 1 : n_threads, body_stream);
    else sg::Pool::get().run(n_threads < 1 ? 1 : n_threads, body);
    SG_CHECK_ARG(bad.load() == 0, "ev_expand_obs: a record holds an out-of-range clock or MOER day");
    for (int j = 0; j < n_ovf; ++j) {                                         // envs with more than 16 active sessions
        const int64_t b = h_ovf_idx[j];
        SG_CHECK_ARG(b >= 0 && b < num_envs, "ev_expand_obs: overflow index out of range");
        std::memcpy(h_obs + b * obs_dim, h_ovf_rows + (int64_t)j * 2 * n, sizeof(float) * 2 * n);
    }
    return 0;
}

}  // extern "C"

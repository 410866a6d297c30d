// GMM day-trace sampler: fills the device-resident trace pool of the EV kernels.
//
// Replaces, for a whole pool of days at once, GMMsTraceGenerator._sample / ._create_events
// (sustaingym/envs/evcharging/event_generation.py:416-515) + the packing the reference does into acnsim
// PluginEvents at every reset (:149-207): one thread samples one day — session count from the empirical
// counts, int(1.2 n) mixture candidates per batch grouped by component (what GaussianMixture.sample
// returns), the reference's bound / ordering filters, rescaling to 5-minute periods and kWh, the cut to
// the first n, the stable sort by arrival and the sequential station draw among the stations free at
// the arrival — and writes the day straight into the layouts ev.cu reads (sessions[D,K,n] 16-byte
// records, arrivals[D,Smax+1] event list), so a reset stays an index bump and a fresh pool never
// touches the host.
//
// The reference's random streams cannot be reproduced (NumPy Generator + scikit-learn RandomState), so
// the sampler defines its own: Philox4x32-10 keyed by the seed, counter = (block, day).  Every double
// operation is individually rounded (__dmul_rn / __dadd_rn / __ddiv_rn / __dsqrt_rn, an explicit
// logarithm) so that oracle/ev_gmm.py reproduces the kernel bit for bit.
#include "common.cuh"

namespace sg {
namespace {

constexpr int kCandMax = 256;     // candidate sessions of one day (Caltech <= 44/day, JPL <= ~160)
constexpr int kCompMax = 64;

struct Philox {
    uint32_t key0, key1, day0, day1, block;
    uint32_t buf[4];
    int have;
    __device__ void refill() {
        uint32_t c0 = block, c1 = day0, c2 = day1, c3 = 0, k0 = key0, k1 = key1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        buf[0] = c0; buf[1] = c1; buf[2] = c2; buf[3] = c3;
        ++block;
        have = 4;
    }
    __device__ uint32_t u32() {
        if (have == 0) refill();
        const uint32_t v = buf[4 - have];
        --have;
        return v;
    }
    __device__ double uniform() {      // 53-bit uniform in [0, 1)
        const uint32_t a = u32() >> 5, b = u32() >> 6;
        return __ddiv_rn(__dadd_rn(__dmul_rn((double)a, 67108864.0), (double)b), 9007199254740992.0);
    }
};

// ln(x): reduction to [sqrt(1/2), sqrt(2)) + atanh series, one rounding per operation (oracle/ev_gmm.py::det_log)
__device__ double det_log(const double x) {
    int e;
    double m = frexp(x, &e);
    if (m < 0.7071067811865476) { m = __dmul_rn(m, 2.0); --e; }
    const double s = __ddiv_rn(__dadd_rn(m, -1.0), __dadd_rn(m, 1.0));
    const double z = __dmul_rn(s, s);
    const double coef[11] = {2.0 / 3, 2.0 / 5, 2.0 / 7, 2.0 / 9, 2.0 / 11, 2.0 / 13, 2.0 / 15, 2.0 / 17, 2.0 / 19,
                             2.0 / 21, 2.0 / 23};
    double p = coef[10];
#pragma unroll
    for (int k = 9; k >= 0; --k) p = __dadd_rn(__dmul_rn(p, z), coef[k]);
    const double r = __dadd_rn(__dmul_rn(2.0, s), __dmul_rn(s, __dmul_rn(z, p)));
    return __dadd_rn(__dmul_rn((double)e, 0.6931471805599453), r);
}

__device__ void normal_pair(Philox& rs, double& z0, double& z1) {      // Marsaglia polar method
    double u1, u2, s;
    do {
        u1 = __dadd_rn(__dmul_rn(2.0, rs.uniform()), -1.0);
        u2 = __dadd_rn(__dmul_rn(2.0, rs.uniform()), -1.0);
        s = __dadd_rn(__dmul_rn(u1, u1), __dmul_rn(u2, u2));
    } while (!(s > 0.0 && s < 1.0));
    const double f = __dsqrt_rn(__ddiv_rn(__dmul_rn(-2.0, det_log(s)), s));
    z0 = __dmul_rn(u1, f);
    z1 = __dmul_rn(u2, f);
}

__global__ void __launch_bounds__(64)
ev_gmm_days_kernel(const sg_ev_gmm g, const long long n_days, const long long first_day, const unsigned long long seed,
                   const double energy_cap, const double kwh_per_step, const int K, const int smax,
                   sg_ev_session* __restrict__ sessions, uint32_t* __restrict__ arrivals,
                   int32_t* __restrict__ dropped_out, double* __restrict__ max_kwh_out) {
    const long long d0 = (long long)blockIdx.x * blockDim.x;
    const long long d = d0 + threadIdx.x;
    const int nc = g.n_components, n_st = g.n_stations;
    {   // the block's days start empty: one contiguous range of each table, filled with coalesced stores
        const long long nd = min((long long)blockDim.x, n_days - d0);
        uint4 e;                                             // {arrival, departure, est_departure = 32767, reserved 0 | energy 0.0}
        e.x = 0x7FFF7FFFu; e.y = 0x00007FFFu; e.z = 0u; e.w = 0u;
        uint4* rec4 = reinterpret_cast<uint4*>(sessions + d0 * K * n_st);
        for (long long i = threadIdx.x; i < nd * K * n_st; i += blockDim.x) rec4[i] = e;
        uint32_t* rows = arrivals + d0 * (smax + 1);
        for (long long i = threadIdx.x; i < nd * (smax + 1); i += blockDim.x) rows[i] = 0x7FFFu;
    }
    __syncthreads();
    if (d >= n_days) return;
    Philox rs;
    const unsigned long long day = (unsigned long long)(first_day + d);
    rs.key0 = (uint32_t)seed; rs.key1 = (uint32_t)(seed >> 32);
    rs.day0 = (uint32_t)day; rs.day1 = (uint32_t)(day >> 32);
    rs.block = 0; rs.have = 0;
    sg_ev_session* rec = sessions + d * K * n_st;

    int n = g.counts[min((int)__dmul_rn(rs.uniform(), (double)g.n_counts), g.n_counts - 1)];
    n = min(n, kCandMax);
    const int m = (int)__dmul_rn((double)n, 1.2);
    short c_arr[kCandMax], c_dep[kCandMax], c_est[kCandMax];
    double c_kwh[kCandMax];
    int ncand = 0;
    for (int batch = 0; ncand < n && batch < 64; ++batch) {
        unsigned char per_comp[kCompMax];
        int extra[kCompMax];                                 // counts beyond 255 (never with nc = 30, m <= 307)
        for (int c = 0; c < nc; ++c) { per_comp[c] = 0; extra[c] = 0; }
        for (int i = 0; i < m; ++i) {
            const double u = rs.uniform();
            int c = 0;
            while (c < nc - 1 && !(g.cum_weights[c] > u)) ++c;
            if (per_comp[c] == 255) ++extra[c]; else ++per_comp[c];
        }
        for (int c = 0; c < nc; ++c) {
            const int cnt = per_comp[c] + extra[c];
            const double* mu = g.means + c * 4;
            const double* L = g.chol + c * 16;
            for (int q = 0; q < cnt; ++q) {
                double z[4], x[4];
                normal_pair(rs, z[0], z[1]);
                normal_pair(rs, z[2], z[3]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    double acc = mu[i];
#pragma unroll
                    for (int j = 0; j <= i; ++j) acc = __dadd_rn(acc, __dmul_rn(L[i * 4 + j], z[j]));
                    x[i] = acc;
                }
                if (!(0.0 <= x[0] && x[1] < 1.0 && x[2] < 1.0 && x[3] >= 0.0)) continue;     // :441-444
                const double fa = floor(__ddiv_rn(__dmul_rn(1440.0, x[0]), 5.0));            // :447-449
                const double fd = floor(__ddiv_rn(__dmul_rn(1440.0, x[1]), 5.0));
                const double fe = floor(__ddiv_rn(__dmul_rn(1440.0, x[2]), 5.0));
                if (!(fa < fd && fa < fe)) continue;                                         // :452-455
                if (ncand < n) {                                                             // :460-463
                    c_arr[ncand] = (short)fa; c_dep[ncand] = (short)fd; c_est[ncand] = (short)fe;
                    c_kwh[ncand] = fmin(fmax(__dmul_rn(x[3], 100.0), 0.0), energy_cap);      // :458, :486
                    ++ncand;
                }
            }
        }
    }
    // stable sort by arrival (:490)
    unsigned char order[kCandMax];
    for (int i = 0; i < ncand; ++i) {
        int p = i;
        while (p > 0 && c_arr[order[p - 1]] > c_arr[i]) { order[p] = order[p - 1]; --p; }
        order[p] = (unsigned char)i;
    }
    // sequential station draw (:492-515)
    short station_dep[SG_EV_MAX_STATIONS];
    unsigned char used[SG_EV_MAX_STATIONS];
    for (int i = 0; i < n_st; ++i) { station_dep[i] = -1; used[i] = 0; }
    uint32_t keys[kCandMax];
    int n_ev = 0, dropped = 0;
    double max_kwh = 0.0;
    for (int q = 0; q < ncand; ++q) {
        const int ci = order[q];
        const int arr = c_arr[ci], dep = c_dep[ci];
        const double u = rs.uniform();
        int n_avail = 0, last = -1;
        double total = 0.0;
        for (int i = 0; i < n_st; ++i)
            if (station_dep[i] < arr) { ++n_avail; last = i; total = __dadd_rn(total, g.station_p[i]); }
        if (n_avail == 0) { ++dropped; continue; }
        int idx = last;
        if (total <= 1e-5) {
            int pick = min((int)__dmul_rn(u, (double)n_avail), n_avail - 1);
            for (int i = 0; i < n_st; ++i)
                if (station_dep[i] < arr && pick-- == 0) { idx = i; break; }
        } else {
            const double target = __dmul_rn(u, total);
            double cum = 0.0;
            for (int i = 0; i < n_st; ++i)
                if (station_dep[i] < arr) {
                    cum = __dadd_rn(cum, g.station_p[i]);
                    if (cum > target) { idx = i; break; }
                }
        }
        if (used[idx] >= K || n_ev >= smax) { ++dropped; continue; }
        station_dep[idx] = (short)max(dep, (int)station_dep[idx]);
        const int slot = used[idx]++;
        sg_ev_session e;
        e.arrival = (int16_t)arr; e.departure = (int16_t)dep; e.est_departure = c_est[ci]; e.reserved = 0;
        e.energy = c_kwh[ci];
        rec[slot * n_st + idx] = e;
        keys[n_ev++] = ((uint32_t)arr << 16) | ((uint32_t)idx << 8) | (uint32_t)slot;
        max_kwh = __dadd_rn(max_kwh, fmin(c_kwh[ci], __dmul_rn((double)(dep - arr), kwh_per_step)));
    }
    // event list sorted by (arrival, station)
    for (int i = 1; i < n_ev; ++i) {
        const uint32_t k = keys[i];
        int p = i;
        while (p > 0 && keys[p - 1] > k) { keys[p] = keys[p - 1]; --p; }
        keys[p] = k;
    }
    uint32_t* row = arrivals + d * (smax + 1);
    for (int i = 0; i < n_ev; ++i) {
        const uint32_t k = keys[i];
        row[i] = (k >> 16) | (((k >> 8) & 0xffu) << 16) | ((k & 0xffu) << 24);
    }
    if (dropped_out) dropped_out[d] = dropped;
    if (max_kwh_out) max_kwh_out[d] = max_kwh;
}

}  // namespace
}  // namespace sg

extern "C" int sg_ev_sample_gmm_days(const sg_ev_gmm* g, int64_t n_days, int64_t first_day, uint64_t seed,
                                     double requested_energy_cap, double kwh_per_step, int32_t max_sessions,
                                     int32_t max_day_sessions, void* sessions, uint32_t* arrivals, int32_t* dropped,
                                     double* max_kwh, void* stream) {
    SG_CHECK_ARG(g != nullptr, "ev_sample_gmm_days: NULL model");
    SG_CHECK_ARG(g->n_components >= 1 && g->n_components <= sg::kCompMax, "ev_sample_gmm_days: n_components must be in [1, 64]");
    SG_CHECK_ARG(g->n_stations >= 1 && g->n_stations <= SG_EV_MAX_STATIONS, "ev_sample_gmm_days: n_stations must be in [1, 64]");
    SG_CHECK_ARG(g->n_counts >= 1, "ev_sample_gmm_days: empty session-count table");
    SG_CHECK_ARG(g->cum_weights && g->means && g->chol && g->counts && g->station_p, "ev_sample_gmm_days: NULL table");
    SG_CHECK_ARG(max_sessions >= 1 && max_sessions <= 254, "ev_sample_gmm_days: max_sessions must be in [1, 254]");
    SG_CHECK_ARG(max_day_sessions >= 1 && max_day_sessions <= sg::kCandMax, "ev_sample_gmm_days: max_day_sessions must be in [1, 256]");
    SG_CHECK_ARG(sessions && arrivals, "ev_sample_gmm_days: NULL output");
    SG_CHECK_ARG(first_day >= 0, "ev_sample_gmm_days: first_day must be >= 0");
    if (n_days <= 0) return 0;
    sg::ev_gmm_days_kernel<<<(unsigned)((n_days + 63) / 64), 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        *g, n_days, first_day, seed, requested_energy_cap, kwh_per_step, max_sessions, max_day_sessions,
        static_cast<sg_ev_session*>(sessions), arrivals, dropped, max_kwh);
    return sg::launch_status("ev_gmm_days_kernel");
}

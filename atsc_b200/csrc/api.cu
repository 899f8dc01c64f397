// api.cu -- context, workspaces, wave orchestration and the C ABI (include/atsc_gpu.h).
//
// There is deliberately NO CPU fallback: every entry point fails with ATSC_ERR_CUDA when no
// CUDA device is usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/atsc_gpu.h"
#include "host_util.h"
#include "kernels.h"

using namespace atsc;

namespace {

constexpr uint64_t WAVE_SAMPLES = 48ull << 20;  // samples per wave (384 MiB of f64)
constexpr uint32_t WAVE_FRAMES = 1u << 18;

struct Device {
    int id = 0;
    cudaStream_t st = nullptr;
    SlotPool pool{};
    double *inv_d2 = nullptr;
    unsigned *queues = nullptr;
    unsigned long long *d_total = nullptr, *h_total = nullptr;
    // geometry cache
    std::map<uint32_t, int> geom_idx;
    std::map<uint32_t, uint32_t> pad_cache;
    std::vector<FftGeom> geoms_host;
    std::vector<void *> geom_allocs;
    FftGeom *geoms_dev = nullptr;
    size_t geoms_dev_cap = 0;
    bool geoms_dirty = false;
    // growable buffers
    FrameWork *d_frames = nullptr, *h_frames = nullptr;
    size_t frames_cap = 0;
    double *d_samples = nullptr;
    size_t samples_cap = 0;
    FftEntry *d_arena = nullptr;
    size_t arena_cap = 0;
    float2 *d_spec_xd = nullptr;  // per-frame half spectra of the wave (k_fft_fwd -> k_fft)
    uint32_t *d_spec_keys = nullptr;
    size_t spec_xd_cap = 0, spec_keys_cap = 0;
    uint8_t *d_payload = nullptr, *h_payload = nullptr;
    size_t payload_cap = 0, h_payload_cap = 0;
    DecFrame *d_dec = nullptr, *h_dec = nullptr;
    size_t dec_cap = 0;
    uint8_t *d_pay_in = nullptr;
    size_t pay_in_cap = 0;
    double *d_out = nullptr;
    size_t out_cap = 0;
    uint32_t *d_status = nullptr, *h_status = nullptr;
    size_t status_cap = 0;
    uint64_t launches = 0;
    // CUDA-event timing of every kernel on this device's stream (ms accumulated since reset):
    // 0 stats, 1 plan+poly, 2 rle, 3 fft, 4 noop+select+scan, 5 emit, 6 decode, 7 unused
    cudaEvent_t ev[10] = {};
    double ms[8] = {};
    std::string err;
};

}  // namespace

struct atsc_ctx {
    std::vector<Device *> devs;
    std::string err;
};

namespace {

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            char b__[512];                                                                         \
            snprintf(b__, sizeof b__, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            D.err = b__;                                                                           \
            return ATSC_ERR_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

template <class T>
int grow(Device &D, T *&p, size_t &cap, size_t need, bool pinned_host = false) {
    if (need <= cap) return ATSC_OK;
    size_t nc = std::max(need, cap + cap / 2);
    if (p) {
        if (pinned_host)
            CK(cudaFreeHost(p));
        else
            CK(cudaFree(p));
        p = nullptr;
    }
    if (pinned_host)
        CK(cudaMallocHost((void **)&p, nc * sizeof(T)));
    else
        CK(cudaMalloc((void **)&p, nc * sizeof(T)));
    cap = nc;
    return ATSC_OK;
}

// ---------------------------------------------------------------- FFT geometry tables
void radices_of(uint32_t len, FftStage *stg, uint32_t *ns) {
    // radices {9, 8, 4, 3, 2}: as few Stockham stages (shared-memory round trips) as possible
    uint8_t rad[16];
    uint32_t k = 0, n = len;
    while (n % 9 == 0) {
        rad[k++] = 9;
        n /= 9;
    }
    while (n % 8 == 0) {
        rad[k++] = 8;
        n /= 8;
    }
    while (n % 4 == 0) {
        rad[k++] = 4;
        n /= 4;
    }
    while (n % 2 == 0) {
        rad[k++] = 2;
        n /= 2;
    }
    while (n % 3 == 0) {
        rad[k++] = 3;
        n /= 3;
    }
    *ns = k;
    uint32_t cur = len, s = 1;
    const uint32_t nw = BLOCK / 32;
    for (uint32_t i = 0; i < k; i++) {
        FftStage &S = stg[i];
        S.r = rad[i];
        S.m = (uint16_t)(cur / rad[i]);
        S.s = (uint16_t)s;
        S.tws = (uint16_t)(len / cur);
        S.nbf = (uint16_t)(len / rad[i]);
        S.dp = (uint16_t)(nw / s);
        S.dq = (uint16_t)(nw % s);
        S.pad = 0;
        S.magic = (65536u + s - 1) / s;
        cur /= rad[i];
        s *= rad[i];
    }
}

template <class T>
int upload_vec(Device &D, const std::vector<T> &v, const T **out) {
    T *p = nullptr;
    CK(cudaMalloc((void **)&p, std::max<size_t>(v.size(), 1) * sizeof(T)));
    D.geom_allocs.push_back(p);
    if (!v.empty()) CK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, D.st));
    CK(cudaStreamSynchronize(D.st));
    *out = p;
    return ATSC_OK;
}

// returns geometry index for transform length L (L must be 2^a 3^b, L >= 2), or <0
int get_geom(Device &D, uint32_t L, int *out_idx) {
    auto it = D.geom_idx.find(L);
    if (it != D.geom_idx.end()) {
        *out_idx = it->second;
        return ATSC_OK;
    }
    if (!atsc_host::is_decomposable(L) || L > (uint32_t)MAX_FFT_LEN) return ATSC_ERR_UNSUPPORTED;
    FftGeom g{};
    g.L = L;
    g.real = (L % 2 == 0) ? 1 : 0;
    g.M = g.real ? L / 2 : L;
    g.Bn = L / 2 + 1;
    // pick M1 x M2: both <= FFT_TLEN, maximise lane utilisation of the 32-wide batches
    double best = -1.0;
    uint32_t b1 = 0, b2 = 0;
    for (uint32_t m1 = 1; m1 <= (uint32_t)FFT_TLEN && m1 <= g.M; m1++) {
        if (g.M % m1) continue;
        uint32_t m2 = g.M / m1;
        if (m2 > (uint32_t)FFT_TLEN) continue;
        double u1 = (double)m1 / (double)(((m1 + 31) / 32) * 32);
        double u2 = (double)m2 / (double)(((m2 + 31) / 32) * 32);
        double score = u1 * u2 - 1e-6 * std::fabs((double)m1 - (double)m2);
        if (score > best) {
            best = score;
            b1 = m1;
            b2 = m2;
        }
    }
    // lengths 2^a 3^7 (every power-of-two frame of 8192..131072 samples): M1 x 243, the split the
    // register-radix forward engine (fft2.cuh) is written for
    const bool fast = g.real && g.M % 243 == 0 &&
                      (g.M / 243 == 288 || g.M / 243 == 144 || g.M / 243 == 72 || g.M / 243 == 36 || g.M / 243 == 18);
    if (fast) {
        b1 = g.M / 243;
        b2 = 243;
    }
    if (!b1) return ATSC_ERR_UNSUPPORTED;
    g.M1 = b1;
    g.M2 = b2;
    radices_of(g.M1, g.st1, &g.ns1);
    radices_of(g.M2, g.st2, &g.ns2);
    const double PI2 = 6.283185307179586476925286766559;
    auto root = [&](double num, double den) {
        double a = PI2 * num / den;
        float2 r;
        r.x = (float)std::cos(a);
        r.y = (float)(-std::sin(a));
        return r;
    };
    std::vector<float2> twL1(g.M1), twL2(g.M2);
    for (uint32_t j = 0; j < g.M1; j++) twL1[j] = root(j, L);
    for (uint32_t j = 0; j < g.M2; j++) twL2[j] = root((double)j * g.M1, L);
    std::vector<float2> twM(g.M), tw1(g.M1), tw2(g.M2), twL(g.M + 1), twA((size_t)g.M1 * 32), twB((size_t)g.M2 * 32);
    for (uint32_t j = 0; j < g.M; j++) twM[j] = root(j, g.M);
    for (uint32_t j = 0; j < g.M1; j++) tw1[j] = root(j, g.M1);
    for (uint32_t j = 0; j < g.M2; j++) tw2[j] = root(j, g.M2);
    for (uint32_t j = 0; j <= g.M; j++) twL[j] = root(j, L);
    for (uint32_t e = 0; e < g.M1; e++)
        for (uint32_t f = 0; f < 32; f++) twA[(size_t)e * 32 + f] = root((double)e * f, g.M);
    for (uint32_t e = 0; e < g.M2; e++)
        for (uint32_t f = 0; f < 32; f++) twB[(size_t)e * 32 + f] = root((double)e * f, g.M);
    int rc;
    if ((rc = upload_vec(D, twM, &g.twM))) return rc;
    if ((rc = upload_vec(D, tw1, &g.tw1))) return rc;
    if ((rc = upload_vec(D, tw2, &g.tw2))) return rc;
    if ((rc = upload_vec(D, twL, &g.twL))) return rc;
    if ((rc = upload_vec(D, twL1, &g.twL1))) return rc;
    if ((rc = upload_vec(D, twL2, &g.twL2))) return rc;
    if ((rc = upload_vec(D, twA, &g.twA))) return rc;
    if ((rc = upload_vec(D, twB, &g.twB))) return rc;
    g.T4 = nullptr;
    if (fast) {
        std::vector<float2> T4((size_t)g.M);
        for (uint32_t n2 = 0; n2 < g.M2; n2++)
            for (uint32_t k1 = 0; k1 < g.M1; k1++) T4[(size_t)n2 * g.M1 + k1] = root((double)(((uint64_t)k1 * n2) % g.M), g.M);
        if ((rc = upload_vec(D, T4, &g.T4))) return rc;
    }
    int idx = (int)D.geoms_host.size();
    D.geoms_host.push_back(g);
    D.geom_idx[L] = idx;
    D.geoms_dirty = true;
    *out_idx = idx;
    return ATSC_OK;
}

// next_size (utils/mod.rs:32-38) memoised per frame length: the search loop costs tens of
// microseconds for 131072 and would otherwise run once per frame
uint32_t padded_len(Device &D, uint32_t len) {
    auto it = D.pad_cache.find(len);
    if (it != D.pad_cache.end()) return it->second;
    uint32_t L = (uint32_t)atsc_host::next_size(len);
    D.pad_cache[len] = L;
    return L;
}

int sync_geoms(Device &D) {
    if (!D.geoms_dirty) return ATSC_OK;
    int rc = grow(D, D.geoms_dev, D.geoms_dev_cap, D.geoms_host.size() + 8);
    if (rc) return rc;
    CK(cudaMemcpyAsync(D.geoms_dev, D.geoms_host.data(), D.geoms_host.size() * sizeof(FftGeom),
                       cudaMemcpyHostToDevice, D.st));
    CK(cudaStreamSynchronize(D.st));
    D.geoms_dirty = false;
    return ATSC_OK;
}

// ---------------------------------------------------------------- device setup
int device_init(Device &D) {
    CK(cudaSetDevice(D.id));
    CK(cudaStreamCreateWithFlags(&D.st, cudaStreamNonBlocking));
    int rc = kernels_init();
    if (rc) {
        D.err = std::string("kernels_init: ") + cudaGetErrorString((cudaError_t)rc);
        return ATSC_ERR_CUDA;
    }
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, D.id));
    SlotPool &P = D.pool;
    P.rle_slots = 2 * sms;
    P.fft_slots = 2 * sms;  // k_fft / k_decode run two 512-thread CTAs per SM
    P.dec_slots = 2 * sms;
    size_t rs = (size_t)P.rle_slots;
    CK(cudaMalloc((void **)&P.rle_k0, rs * MAX_FRAME * 8));
    CK(cudaMalloc((void **)&P.rle_k1, rs * MAX_FRAME * 8));
    CK(cudaMalloc((void **)&P.rle_i0, rs * MAX_FRAME * 4));
    CK(cudaMalloc((void **)&P.rle_i1, rs * MAX_FRAME * 4));
    CK(cudaMalloc((void **)&P.rle_bnd, rs * (MAX_FRAME + 8) * 4));
    size_t fs = (size_t)P.fft_slots, hb = MAX_FFT_LEN / 2 + 8;
    CK(cudaMalloc((void **)&P.fft_W, fs * MAX_FFT_LEN * sizeof(float2)));
    CK(cudaMalloc((void **)&P.fft_Xd, fs * hb * sizeof(float2)));
    CK(cudaMalloc((void **)&P.fft_keys, fs * hb * 4));
    CK(cudaMalloc((void **)&P.fft_rank, fs * hb * 4));
    CK(cudaMalloc((void **)&P.fft_locD, fs * FFT_DEC_KCAP * 4));
    CK(cudaMalloc((void **)&P.fft_locM, fs * FFT_DEC_KCAP * 4));
    CK(cudaMalloc((void **)&P.fft_ovr, fs * FFT_DEC_KCAP * 4));
    CK(cudaMalloc((void **)&P.fft_cD, fs * FFT_DEC_KCAP * sizeof(float2)));
    CK(cudaMalloc((void **)&P.fft_cM, fs * FFT_DEC_KCAP * sizeof(float2)));
    CK(cudaMalloc((void **)&P.fft_dlist, fs * FFT_DEC_KCAP * sizeof(FftEntry)));
    P.poly_slots = 2 * sms;
    CK(cudaMalloc((void **)&P.poly_slope, (size_t)P.poly_slots * (MAX_FRAME + 8) * 8));
    size_t dsl = (size_t)P.dec_slots;
    CK(cudaMalloc((void **)&P.dec_pts, dsl * (MAX_FRAME + 8) * 8));
    CK(cudaMalloc((void **)&P.dec_mark, dsl * (MAX_FRAME + 8) * 4));
    CK(cudaMalloc((void **)&P.dec_idx, dsl * (MAX_FRAME + 8) * 4));
    CK(cudaMalloc((void **)&D.inv_d2, (size_t)(MAX_FRAME + 8) * 8));
    launch_inv_d2(D.inv_d2, MAX_FRAME + 8, D.st);
    D.launches++;
    for (auto &e : D.ev) CK(cudaEventCreate(&e));
    CK(cudaMalloc((void **)&D.queues, 64 * sizeof(unsigned)));
    CK(cudaMalloc((void **)&D.d_total, 8));
    CK(cudaMallocHost((void **)&D.h_total, 8));
    CK(cudaStreamSynchronize(D.st));
    return ATSC_OK;
}

void device_free(Device &D) {
    cudaSetDevice(D.id);
    SlotPool &P = D.pool;
    void *ptrs[] = {P.rle_k0, P.rle_k1, P.rle_i0, P.rle_i1, P.rle_bnd, P.fft_W, P.fft_Xd, P.fft_keys, P.fft_rank,
                    P.fft_locD, P.fft_locM, P.fft_ovr, P.fft_cD, P.fft_cM, P.fft_dlist, P.poly_slope, P.dec_pts, P.dec_mark,
                    P.dec_idx, D.inv_d2, D.queues, D.d_total, D.geoms_dev, D.d_frames, D.d_samples, D.d_arena,
                    D.d_payload, D.d_dec, D.d_pay_in, D.d_out, D.d_status, D.d_spec_xd, D.d_spec_keys};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (void *p : D.geom_allocs) cudaFree(p);
    void *hp[] = {D.h_total, D.h_frames, D.h_payload, D.h_dec, D.h_status};
    for (void *p : hp)
        if (p) cudaFreeHost(p);
    for (auto &e : D.ev)
        if (e) cudaEventDestroy(e);
    if (D.st) cudaStreamDestroy(D.st);
}

bool is_device_ptr(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// ---------------------------------------------------------------- compress
struct FrameReq {
    uint64_t off;
    uint32_t len;
    uint8_t comp, bounded, select_only, forced;
};

// runs the pipeline for reqs[0..n) whose samples live at d_samples (device); results are left
// in D.h_frames[0..n); payload bytes (if any) in D.h_payload[0..*payload_total)
// direct_dst != nullptr: page-locked destination (capacity direct_cap) the payload is copied to
// straight from the device; *direct is set when that happened (else the bytes are in D.h_payload)
int run_wave(Device &D, const double *d_samples, const std::vector<FrameReq> &reqs, float max_error_f32,
             uint64_t *payload_total, uint8_t *direct_dst = nullptr, uint64_t direct_cap = 0, bool *direct = nullptr) {
    if (direct) *direct = false;
    const uint32_t n = (uint32_t)reqs.size();
    const double max_err = (double)max_error_f32;  // `max_error as f64` (frame/mod.rs:67,87)
    int rc;
    size_t hcap = D.frames_cap;
    if ((rc = grow(D, D.d_frames, D.frames_cap, n))) return rc;
    if ((rc = grow(D, D.h_frames, hcap, D.frames_cap, true))) return rc;
    uint64_t arena = 0, spec = 0;
    bool any_noop = false;
    for (uint32_t i = 0; i < n; i++) {
        FrameWork &f = D.h_frames[i];
        memset(&f, 0, sizeof f);
        const FrameReq &r = reqs[i];
        f.off = r.off;
        f.len = r.len;
        f.comp = r.comp;
        f.bounded = r.bounded;
        f.select_only = r.select_only;
        f.forced = r.forced;
        f.geom = -1;
        f.spec_off = ~0ull;
        uint8_t eff = (r.comp == C_AUTO && r.forced != 0xFF) ? r.forced : r.comp;
        any_noop |= r.comp == C_NOOP;
        if (eff == C_FFT || eff == C_AUTO) {
            uint32_t L = (r.bounded && r.len >= 128) ? padded_len(D, r.len) : r.len;
            if (r.len >= 128) {
                int gi;
                if ((rc = get_geom(D, L, &gi))) {
                    D.err = "FFT length is not of the form 2^a*3^b (unbounded Compressor::FFT needs a decomposable frame length)";
                    return rc;
                }
                f.geom = gi;
                if (D.geoms_host[gi].T4) {
                    f.spec_off = spec;
                    spec += D.geoms_host[gi].M + 8;
                }
            }
            uint32_t mf = std::max<uint32_t>(3, r.len / 100);
            uint32_t kmax = r.bounded ? mf + 17 * std::max<uint32_t>(mf / 2, 1) + 5 * std::max<uint32_t>(mf / 10, 1) : mf;
            uint32_t cap = std::min<uint32_t>(std::min<uint32_t>(kmax, L / 2 + 1), FFT_KCAP);
            f.fft_list_off = arena;
            f.fft_list_cap = cap;
            arena += cap;
        }
    }
    if ((rc = sync_geoms(D))) return rc;
    if ((rc = grow(D, D.d_arena, D.arena_cap, (size_t)arena + 1))) return rc;
    if ((rc = grow(D, D.d_spec_xd, D.spec_xd_cap, (size_t)spec + 1))) return rc;
    if ((rc = grow(D, D.d_spec_keys, D.spec_keys_cap, (size_t)spec + 1))) return rc;
    CK(cudaMemcpyAsync(D.d_frames, D.h_frames, (size_t)n * sizeof(FrameWork), cudaMemcpyHostToDevice, D.st));
    CK(cudaMemsetAsync(D.queues, 0, 64 * sizeof(unsigned), D.st));
    CK(cudaEventRecord(D.ev[0], D.st));
    launch_stats(D.d_frames, n, d_samples, D.queues + 0, D.st);
    CK(cudaEventRecord(D.ev[1], D.st));
    launch_plan(D.d_frames, n, D.st);
    launch_poly(D.d_frames, n, d_samples, max_err, D.inv_d2, D.pool, D.queues + 1, D.st);
    CK(cudaEventRecord(D.ev[2], D.st));
    launch_rle(D.d_frames, n, d_samples, max_err, D.pool, D.queues + 2, D.st);
    CK(cudaEventRecord(D.ev[3], D.st));
    if (spec) {
        launch_fft_fwd(D.d_frames, n, d_samples, max_err, D.geoms_dev, D.pool, D.d_spec_xd, D.d_spec_keys, D.queues + 7, D.st);
        D.launches++;
    }
    launch_fft(D.d_frames, n, d_samples, max_err, D.geoms_dev, D.pool, D.d_arena, D.d_spec_xd, D.d_spec_keys, D.queues + 3, D.st);
    CK(cudaEventRecord(D.ev[4], D.st));
    D.launches += 5;
    if (any_noop) {
        launch_noop_size(D.d_frames, n, d_samples, D.queues + 4, D.st);
        D.launches++;
    }
    launch_select(D.d_frames, n, max_err, D.st);
    launch_scan(D.d_frames, n, D.d_total, D.st);
    D.launches += 2;
    CK(cudaEventRecord(D.ev[5], D.st));
    CK(cudaMemcpyAsync(D.h_total, D.d_total, 8, cudaMemcpyDeviceToHost, D.st));
    CK(cudaStreamSynchronize(D.st));
    CK(cudaGetLastError());
    for (int k = 0; k < 5; k++) {
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, D.ev[k], D.ev[k + 1]));
        D.ms[k] += t;
    }
    uint64_t total = *D.h_total;
    *payload_total = total;
    if (total) {
        const bool to_user = direct_dst && total <= direct_cap;
        if ((rc = grow(D, D.d_payload, D.payload_cap, (size_t)total + 16))) return rc;
        if (!to_user && (rc = grow(D, D.h_payload, D.h_payload_cap, (size_t)total + 16, true))) return rc;
        CK(cudaEventRecord(D.ev[6], D.st));
        launch_emit(D.d_frames, n, d_samples, D.geoms_dev, D.pool, D.d_arena, D.d_payload, D.queues + 5, D.st);
        CK(cudaEventRecord(D.ev[7], D.st));
        D.launches++;
        CK(cudaMemcpyAsync(to_user ? direct_dst : D.h_payload, D.d_payload, (size_t)total, cudaMemcpyDeviceToHost, D.st));
        if (direct) *direct = to_user;
    }
    CK(cudaMemcpyAsync(D.h_frames, D.d_frames, (size_t)n * sizeof(FrameWork), cudaMemcpyDeviceToHost, D.st));
    CK(cudaStreamSynchronize(D.st));
    CK(cudaGetLastError());
    if (total) {
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, D.ev[6], D.ev[7]));
        D.ms[5] += t;
    }
    return ATSC_OK;
}

void fill_out(const FrameWork &f, atsc_frame_out &o, uint64_t payload_base) {
    o.compressor = f.winner;
    o.near_tie = f.near_tie;
    o.iterations = f.iterations;
    o.payload_len = f.payload_len;
    o.payload_off = payload_base + f.payload_off;
    o.error = f.error;
    o.cand_error[0] = f.fft_valid ? f.fft_err : 0.0;
    o.cand_error[1] = f.poly_valid ? f.poly_err : 0.0;
    o.cand_error[2] = 0.0;
    o.cand_size[0] = f.fft_valid == 1 ? f.fft_size : 0;
    o.cand_size[1] = f.poly_valid == 1 ? f.poly_size : 0;
    o.cand_size[2] = f.rle_valid == 1 ? f.rle_size : 0;
    o.reserved = 0;
}

struct PayloadSink {
    uint8_t *buf;
    uint64_t cap, used;
    bool overflow;
    bool pinned;  // buf is page-locked host memory: payloads are copied to it straight from the device
};

bool is_pinned_host(const void *p) {
    cudaPointerAttributes a;
    if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

static const uint32_t COMPRESSION_SPEED[7] = {2147483647u, 4096, 2048, 1024, 512, 256, 128};  // frame/mod.rs:22

// compress frames idx[0..m) on device D
int compress_on_device(Device &D, const double *samples, bool dev_ptr, const uint64_t *frame_off,
                       const uint32_t *frame_len, const uint32_t *idx, uint32_t m, uint8_t compressor,
                       float max_error, uint32_t speed, int bounded, atsc_frame_out *out, PayloadSink &sink) {
    CK(cudaSetDevice(D.id));
    const bool sampled = bounded && compressor == C_AUTO && speed > 0;
    const uint32_t sample_n = COMPRESSION_SPEED[speed];
    uint32_t pos = 0;
    std::vector<FrameReq> reqs, sel;
    std::vector<uint32_t> sel_of;
    while (pos < m) {
        // ---- cut a wave
        uint64_t tot = 0;
        uint32_t end = pos;
        uint64_t lo = ~0ull, hi = 0;
        while (end < m && end - pos < WAVE_FRAMES) {
            uint32_t fi = idx[end];
            if (tot && tot + frame_len[fi] > WAVE_SAMPLES) break;
            tot += frame_len[fi];
            lo = std::min(lo, frame_off[fi]);
            hi = std::max(hi, frame_off[fi] + frame_len[fi]);
            end++;
        }
        const uint32_t n = end - pos;
        // ---- samples on the device
        const double *d_samples = samples;
        bool packed = false;
        if (!dev_ptr) {
            uint64_t span = hi - lo;
            packed = span > 2 * tot + 4096;
            int rc = grow(D, D.d_samples, D.samples_cap, (size_t)(packed ? tot : span) + 8);
            if (rc) return rc;
            if (!packed) {
                CK(cudaMemcpyAsync(D.d_samples, samples + lo, span * 8, cudaMemcpyHostToDevice, D.st));
            } else {
                uint64_t o = 0;
                for (uint32_t k = pos; k < end; k++) {
                    uint32_t fi = idx[k];
                    CK(cudaMemcpyAsync(D.d_samples + o, samples + frame_off[fi], (size_t)frame_len[fi] * 8,
                                       cudaMemcpyHostToDevice, D.st));
                    o += frame_len[fi];
                }
            }
            d_samples = D.d_samples;
        }
        reqs.clear();
        uint64_t po = 0;
        for (uint32_t k = pos; k < end; k++) {
            uint32_t fi = idx[k];
            FrameReq r;
            r.off = dev_ptr ? frame_off[fi] : (packed ? po : frame_off[fi] - lo);
            po += frame_len[fi];
            r.len = frame_len[fi];
            r.comp = compressor;
            r.bounded = bounded ? 1 : 0;
            r.select_only = 0;
            r.forced = 0xFF;
            reqs.push_back(r);
        }
        std::vector<uint8_t> tie1(n, 0);
        std::vector<atsc_frame_out> diag1;
        if (sampled) {
            // frame/mod.rs:89-111: pick the compressor on data[0..sample], then compress everything with it
            sel.clear();
            sel_of.clear();
            for (uint32_t k = 0; k < n; k++) {
                if (reqs[k].len >= sample_n) {
                    FrameReq r = reqs[k];
                    r.len = sample_n;
                    r.select_only = 1;
                    sel.push_back(r);
                    sel_of.push_back(k);
                }
            }
            if (!sel.empty()) {
                uint64_t pt;
                int rc = run_wave(D, d_samples, sel, max_error, &pt);
                if (rc) return rc;
                diag1.resize(n);
                for (size_t s = 0; s < sel.size(); s++) {
                    const FrameWork &f = D.h_frames[s];
                    reqs[sel_of[s]].forced = f.winner;
                    tie1[sel_of[s]] = f.near_tie;
                    fill_out(f, diag1[sel_of[s]], 0);
                }
            }
        }
        uint64_t ptotal;
        bool direct = false;
        int rc = run_wave(D, d_samples, reqs, max_error, &ptotal, sink.pinned ? sink.buf + sink.used : nullptr,
                          sink.pinned && sink.cap > sink.used ? sink.cap - sink.used : 0, &direct);
        if (rc) return rc;
        for (uint32_t k = 0; k < n; k++) {
            const FrameWork &f = D.h_frames[k];
            atsc_frame_out &o = out[idx[pos + k]];
            fill_out(f, o, sink.used);
            if (sampled && reqs[k].forced != 0xFF && !f.is_const) {
                o.near_tie |= tie1[k];
                for (int c = 0; c < 3; c++) {
                    o.cand_error[c] = diag1[k].cand_error[c];
                    o.cand_size[c] = diag1[k].cand_size[c];
                }
            }
        }
        if (sink.used + ptotal > sink.cap)
            sink.overflow = true;
        else if (ptotal && !direct)
            memcpy(sink.buf + sink.used, D.h_payload, ptotal);
        sink.used += ptotal;
        pos = end;
    }
    return ATSC_OK;
}

// ---------------------------------------------------------------- decompress
int decompress_on_device(Device &D, const atsc_frame_in *frames, const uint32_t *idx, uint32_t m,
                         const uint8_t *payloads, uint64_t payload_bytes, double *out, bool out_dev) {
    CK(cudaSetDevice(D.id));
    uint32_t pos = 0;
    while (pos < m) {
        uint64_t tot = 0;
        uint32_t end = pos;
        uint64_t plo = ~0ull, phi = 0;
        while (end < m && end - pos < WAVE_FRAMES) {
            const atsc_frame_in &f = frames[idx[end]];
            if (tot && tot + f.sample_count > WAVE_SAMPLES) break;
            tot += f.sample_count;
            plo = std::min<uint64_t>(plo, f.payload_off);
            phi = std::max<uint64_t>(phi, f.payload_off + f.payload_len);
            end++;
        }
        const uint32_t n = end - pos;
        if (phi > payload_bytes) {
            D.err = "frame payload range exceeds payload_bytes";
            return ATSC_ERR_ARG;
        }
        int rc;
        size_t hc = D.dec_cap;
        if ((rc = grow(D, D.d_dec, D.dec_cap, n))) return rc;
        if ((rc = grow(D, D.h_dec, hc, D.dec_cap, true))) return rc;
        hc = D.status_cap;
        if ((rc = grow(D, D.d_status, D.status_cap, n))) return rc;
        if ((rc = grow(D, D.h_status, hc, D.status_cap, true))) return rc;
        if ((rc = grow(D, D.d_pay_in, D.pay_in_cap, (size_t)(phi - plo) + 64))) return rc;
        if (!out_dev && (rc = grow(D, D.d_out, D.out_cap, (size_t)tot + 8))) return rc;
        uint64_t oo = 0;
        for (uint32_t k = 0; k < n; k++) {
            const atsc_frame_in &f = frames[idx[pos + k]];
            DecFrame &d = D.h_dec[k];
            memset(&d, 0, sizeof d);
            d.payload_off = f.payload_off - plo;
            d.payload_len = f.payload_len;
            d.sample_count = f.sample_count;
            d.out_off = out_dev ? f.out_off : oo;
            oo += f.sample_count;
            d.comp = f.compressor;
            d.geom = -1;
            if (f.sample_count == 0 || f.sample_count > (uint32_t)MAX_FRAME) {
                D.err = "frame sample_count out of range (1..131072)";
                return ATSC_ERR_ARG;
            }
            if (f.compressor == C_FFT && f.sample_count >= 128) {
                int gi;
                if ((rc = get_geom(D, padded_len(D, f.sample_count), &gi))) return rc;
                d.geom = gi;
            }
        }
        if ((rc = sync_geoms(D))) return rc;
        CK(cudaMemcpyAsync(D.d_dec, D.h_dec, (size_t)n * sizeof(DecFrame), cudaMemcpyHostToDevice, D.st));
        CK(cudaMemcpyAsync(D.d_pay_in, payloads + plo, (size_t)(phi - plo), cudaMemcpyHostToDevice, D.st));
        CK(cudaMemsetAsync(D.queues, 0, 64 * sizeof(unsigned), D.st));
        double *d_out = out_dev ? out : D.d_out;
        CK(cudaEventRecord(D.ev[8], D.st));
        launch_decode(D.d_dec, n, D.d_pay_in, d_out, D.geoms_dev, D.pool, D.inv_d2, D.d_status, D.queues + 6, D.st);
        CK(cudaEventRecord(D.ev[9], D.st));
        D.launches++;
        CK(cudaMemcpyAsync(D.h_status, D.d_status, (size_t)n * 4, cudaMemcpyDeviceToHost, D.st));
        if (!out_dev) {
            // coalesce frames that are adjacent in the caller's output
            uint32_t k = 0;
            uint64_t src = 0;
            while (k < n) {
                const atsc_frame_in &f0 = frames[idx[pos + k]];
                uint64_t dst = f0.out_off, len = f0.sample_count;
                uint32_t j = k + 1;
                while (j < n && frames[idx[pos + j]].out_off == dst + len) {
                    len += frames[idx[pos + j]].sample_count;
                    j++;
                }
                CK(cudaMemcpyAsync(out + dst, D.d_out + src, len * 8, cudaMemcpyDeviceToHost, D.st));
                src += len;
                k = j;
            }
        }
        CK(cudaStreamSynchronize(D.st));
        CK(cudaGetLastError());
        {
            float t = 0.f;
            CK(cudaEventElapsedTime(&t, D.ev[8], D.ev[9]));
            D.ms[6] += t;
        }
        for (uint32_t k = 0; k < n; k++)
            if (D.h_status[k]) {
                char b[128];
                snprintf(b, sizeof b, "frame %u: malformed or unsupported payload (code %u)", idx[pos + k], D.h_status[k]);
                D.err = b;
                return D.h_status[k] == 4 ? ATSC_ERR_UNSUPPORTED : ATSC_ERR_FORMAT;
            }
        pos = end;
    }
    return ATSC_OK;
}

// contiguous ranges of frames balanced by sample count (atsc_plan_shards, ingest.cpp)
std::vector<std::vector<uint32_t>> shard(const uint32_t *lens, uint32_t n, size_t ndev) {
    std::vector<std::vector<uint32_t>> parts(ndev);
    std::vector<uint32_t> first(ndev + 1);
    atsc_plan_shards(lens, n, (uint32_t)ndev, first.data());
    for (size_t d = 0; d < ndev; d++)
        for (uint32_t i = first[d]; i < first[d + 1]; i++) parts[d].push_back(i);
    return parts;
}

}  // namespace

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

int atsc_gpu_create(const int *device_ids, int n_devices, atsc_ctx **out) {
    if (!out) return ATSC_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return ATSC_ERR_CUDA;  // no CPU fallback
    atsc_ctx *ctx = new atsc_ctx();
    int def = 0;
    if (!device_ids || n_devices <= 0) {
        device_ids = &def;
        n_devices = 1;
    }
    for (int i = 0; i < n_devices; i++) {
        if (device_ids[i] < 0 || device_ids[i] >= count) {
            atsc_gpu_destroy(ctx);
            return ATSC_ERR_ARG;
        }
        Device *D = new Device();
        D->id = device_ids[i];
        ctx->devs.push_back(D);
        int rc = device_init(*D);
        if (rc) {
            fprintf(stderr, "atsc_gpu_create: %s\n", D->err.c_str());
            atsc_gpu_destroy(ctx);
            return rc;
        }
    }
    *out = ctx;
    return ATSC_OK;
}

void atsc_gpu_destroy(atsc_ctx *ctx) {
    if (!ctx) return;
    for (Device *D : ctx->devs) {
        device_free(*D);
        delete D;
    }
    delete ctx;
}

const char *atsc_gpu_last_error(const atsc_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

void *atsc_gpu_host_alloc(uint64_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
void atsc_gpu_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

uint64_t atsc_gpu_launch_count(const atsc_ctx *ctx) {
    uint64_t n = 0;
    if (ctx)
        for (Device *D : ctx->devs) n += D->launches;
    return n;
}

void atsc_gpu_kernel_ms(atsc_ctx *ctx, double *out8, int reset) {
    for (int k = 0; k < 8; k++) out8[k] = 0.0;
    if (!ctx) return;
    for (Device *D : ctx->devs)
        for (int k = 0; k < 8; k++) {
            out8[k] += D->ms[k];
            if (reset) D->ms[k] = 0.0;
        }
}

int atsc_gpu_compress_frames(atsc_ctx *ctx, const double *samples, const uint64_t *frame_off,
                             const uint32_t *frame_len, uint32_t n_frames, uint8_t compressor, float max_error,
                             uint32_t speed, int bounded, atsc_frame_out *out, uint8_t *payload_buf,
                             uint64_t payload_cap, uint64_t *payload_used) {
    if (!ctx) return ATSC_ERR_ARG;
    if (payload_used) *payload_used = 0;
    if (n_frames == 0) return ATSC_OK;
    if (!samples || !frame_off || !frame_len || !out || (!payload_buf && payload_cap)) {
        ctx->err = "null argument";
        return ATSC_ERR_ARG;
    }
    if (compressor > 6 || speed > 6 || !(max_error >= 0.0f)) {
        ctx->err = "compressor must be 0..6, speed 0..6, max_error >= 0";
        return ATSC_ERR_ARG;
    }
    if (compressor == ATSC_AUTO && !bounded) {
        ctx->err = "Compressor::Auto has no unbounded compress (reference: todo!())";
        return ATSC_ERR_UNSUPPORTED;
    }
    for (uint32_t i = 0; i < n_frames; i++)
        if (frame_len[i] == 0 || frame_len[i] > (uint32_t)MAX_FRAME) {
            ctx->err = "frame_len must be 1..131072 (optimizer/mod.rs:27)";
            return ATSC_ERR_ARG;
        }
    const bool dev_ptr = is_device_ptr(samples);
    const size_t nd = dev_ptr ? 1 : ctx->devs.size();
    if (nd == 1) {
        std::vector<uint32_t> idx(n_frames);
        for (uint32_t i = 0; i < n_frames; i++) idx[i] = i;
        PayloadSink sink{payload_buf, payload_cap, 0, false, is_pinned_host(payload_buf)};
        Device &D = *ctx->devs[0];
        int rc = compress_on_device(D, samples, dev_ptr, frame_off, frame_len, idx.data(), n_frames, compressor,
                                    max_error, speed, bounded, out, sink);
        if (rc) {
            ctx->err = D.err;
            return rc;
        }
        if (payload_used) *payload_used = sink.used;
        if (sink.overflow) {
            ctx->err = "payload_buf too small";
            return ATSC_ERR_CAPACITY;
        }
        return ATSC_OK;
    }
    // several devices: contiguous frame ranges, one host thread per device, no collective
    auto parts = shard(frame_len, n_frames, nd);
    std::vector<std::vector<uint8_t>> bufs(nd);
    std::vector<PayloadSink> sinks(nd);
    std::vector<int> rcs(nd, 0);
    std::vector<std::thread> th;
    for (size_t d = 0; d < nd; d++) {
        uint64_t cap = 0;
        for (uint32_t i : parts[d]) cap += (uint64_t)frame_len[i] * 16 + 64;  // worst case: RLE of all-distinct f64
        bufs[d].resize(cap);
        sinks[d] = PayloadSink{bufs[d].data(), cap, 0, false, false};
        th.emplace_back([&, d]() {
            if (parts[d].empty()) return;
            rcs[d] = compress_on_device(*ctx->devs[d], samples, false, frame_off, frame_len, parts[d].data(),
                                        (uint32_t)parts[d].size(), compressor, max_error, speed, bounded, out, sinks[d]);
        });
    }
    for (auto &t : th) t.join();
    uint64_t used = 0;
    bool overflow = false;
    for (size_t d = 0; d < nd; d++) {
        if (rcs[d]) {
            ctx->err = ctx->devs[d]->err;
            return rcs[d];
        }
        for (uint32_t i : parts[d]) out[i].payload_off += used;
        if (used + sinks[d].used > payload_cap)
            overflow = true;
        else if (sinks[d].used)
            memcpy(payload_buf + used, bufs[d].data(), sinks[d].used);
        used += sinks[d].used;
    }
    if (payload_used) *payload_used = used;
    if (overflow) {
        ctx->err = "payload_buf too small";
        return ATSC_ERR_CAPACITY;
    }
    return ATSC_OK;
}

int atsc_gpu_decompress_frames(atsc_ctx *ctx, const atsc_frame_in *frames, uint32_t n_frames,
                               const uint8_t *payloads, uint64_t payload_bytes, double *out_samples) {
    if (!ctx) return ATSC_ERR_ARG;
    if (n_frames == 0) return ATSC_OK;
    if (!frames || !payloads || !out_samples) {
        ctx->err = "null argument";
        return ATSC_ERR_ARG;
    }
    for (uint32_t i = 0; i < n_frames; i++)
        if (frames[i].compressor > 6 || frames[i].compressor == ATSC_AUTO) {
            ctx->err = "frame compressor must be a concrete compressor (reference: todo!())";
            return ATSC_ERR_UNSUPPORTED;
        }
    const bool out_dev = is_device_ptr(out_samples);
    const size_t nd = out_dev ? 1 : ctx->devs.size();
    std::vector<uint32_t> lens(n_frames);
    for (uint32_t i = 0; i < n_frames; i++) lens[i] = frames[i].sample_count;
    auto parts = shard(lens.data(), n_frames, nd);
    std::vector<int> rcs(nd, 0);
    if (nd == 1) {
        rcs[0] = decompress_on_device(*ctx->devs[0], frames, parts[0].data(), n_frames, payloads, payload_bytes,
                                      out_samples, out_dev);
    } else {
        std::vector<std::thread> th;
        for (size_t d = 0; d < nd; d++)
            th.emplace_back([&, d]() {
                if (parts[d].empty()) return;
                rcs[d] = decompress_on_device(*ctx->devs[d], frames, parts[d].data(), (uint32_t)parts[d].size(),
                                              payloads, payload_bytes, out_samples, false);
            });
        for (auto &t : th) t.join();
    }
    for (size_t d = 0; d < nd; d++)
        if (rcs[d]) {
            ctx->err = ctx->devs[d]->err;
            return rcs[d];
        }
    return ATSC_OK;
}

}  // extern "C"

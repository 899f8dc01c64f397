// api.cu -- context, workspaces, wave orchestration and the C ABI (include/atsc_gpu.h).
//
// There is deliberately NO CPU fallback: every entry point fails with ATSC_ERR_CUDA when no
// CUDA device is usable.
//
// A call's frames are cut into waves; each wave is one asynchronous sequence
//   [H2D samples] -> stats -> plan -> poly -> rle -> fft_fwd -> fft -> select -> scan -> emit -> D2H records
// on the stream of one of the device's ENGINES (stream + private workspaces).  Several waves are
// in flight at once, so one wave's copies and kernel tails overlap the next wave's kernels.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
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

constexpr uint32_t WAVE_FRAMES = 1u << 18;
constexpr int MAX_ENGINES = 8;
constexpr size_t GEOM_CAP = 512;  // distinct transform lengths 2^a 3^b <= 139968 (about 120 exist)

struct FrameReq {
    uint64_t off;
    uint32_t len;
    uint8_t comp, bounded, select_only, forced;
};

// device-side control words of one wave
struct WaveCtl {
    unsigned long long total;  // payload bytes of the wave (k_scan)
    unsigned overflow;         // k_emit: total did not fit the payload buffer
    unsigned pad;
};

// a wave that has been issued on an engine and not yet collected
struct WaveJob {
    bool active = false;
    uint32_t pos = 0, n = 0;          // frames idx[pos .. pos+n) of the call
    const double *d_samples = nullptr;
    std::vector<FrameReq> reqs;
    std::vector<uint8_t> tie1;        // sampled selection pass: its near-tie bits / diagnostics
    std::vector<atsc_frame_out> diag1;
    bool sampled = false, any_emit_ev = false;
};

struct Engine {
    cudaStream_t st = nullptr;
    bool ready = false;  // every workspace allocated (a partly set-up engine is torn down, never used)
    SlotPool pool{};
    unsigned *queues = nullptr;
    WaveCtl *d_ctl = nullptr, *h_ctl = nullptr;
    FrameWork *d_frames = nullptr, *h_frames = nullptr;
    size_t frames_cap = 0;
    ChunkRef *d_chunks = nullptr, *h_chunks = nullptr;  // stats work items (stats.cuh)
    StatsPart *d_parts = nullptr;
    size_t chunks_cap = 0, parts_cap = 0;
    double *d_samples = nullptr;
    size_t samples_cap = 0;
    FftEntry *d_arena = nullptr;
    size_t arena_cap = 0;
    uint32_t *d_items = nullptr, *h_items = nullptr;  // frames of the wave that go through k_front, or (k_sfold) k_probe's list
    size_t items_cap = 0;
    float4 *d_fold = nullptr;  // k_sfold's probe folds, SF_FOLD_SLOTS per frame (sfold.cuh)
    size_t fold_cap = 0;
    std::vector<uint8_t> fronted;  // issue_wave scratch
    ChunkRef *d_pitems = nullptr, *h_pitems = nullptr;  // (frame, part) work items of the queue-driven k_poly1 (ATSC_POLY1_STATIC=0)
    double *d_ppart = nullptr;                          // their partial MAPE sums (P1_PARTS per item with k_poly1s)
    P1Item *d_p1list = nullptr;                         // k_plan's compacted item descriptors for k_poly1s
    size_t pitems_cap = 0, ppart_cap = 0, p1list_cap = 0;
    float2 *d_spec_xd = nullptr;  // per-frame half spectra of the wave (k_fft_fwd -> k_fft)
    uint32_t *d_spec_keys = nullptr;
    size_t spec_xd_cap = 0, spec_keys_cap = 0;
    uint8_t *d_payload = nullptr;
    size_t payload_cap = 0;
    // decode
    DecFrame *d_dec = nullptr, *h_dec = nullptr;
    size_t dec_cap = 0;
    uint8_t *d_pay_in = nullptr;
    size_t pay_in_cap = 0;
    double *d_out = nullptr;
    size_t out_cap = 0;
    uint32_t *d_status = nullptr, *h_status = nullptr;
    size_t status_cap = 0;
    // CUDA events around every kernel of the wave (kernel_ms)
    cudaEvent_t ev[14] = {};  // 0-7 compress pipeline, 8-9 decode, 10-11 between the FFT kernels, 12-13 front end (k_sfold / k_front)
    WaveJob job;
    bool dec_active = false;  // a decode wave is pending on this engine
    uint32_t dec_pos = 0, dec_n = 0;
};

struct Device {
    int id = 0;
    cudaStream_t st = nullptr;  // setup stream (tables)
    int n_engines = 0, sms = 0;
    // Front end of the big frames (ATSC_FRONT):
    //   2 (default): k_sfold (sfold.cuh) reads the big Auto frames once for the stats AND stage 1 of the FFT
    //      probe; the Polynomial candidate stays with k_poly.  +1.8 % on the bench fleet over the separate passes.
    //   1: k_front (front.cuh) also evaluates the first Polynomial step in that single read.  On B200 the fused
    //      kernel is issue bound and, holding an SM's whole shared memory, cannot overlap the other engines' waves:
    //      8.16 vs 6.36 ms per bench step, so it is opt-in.
    //   0: separate passes (k_stats, k_poly, k_fft_fwd's own probe).
    int front = 2;
    bool probe_kernel = true;  // ATSC_PROBE_KERNEL=0: k_fft_fwd runs the probe tails of k_sfold's frames itself
    bool poly_items = true;  // ATSC_POLY_ITEMS=0: k_poly evaluates the first step of the big frames itself
    int poly_items_maxf = 3;  // ATSC_POLY_ITEMS_MAXF: items only while the big frames number at most this many per k_poly CTA slot
    bool poly1_static = true;  // ATSC_POLY1_STATIC=0: the queue-driven k_poly1 instead of k_poly1_prep + k_poly1s (A/B runs)
    bool front_poly = true;  // ... including the first Polynomial step (ATSC_FRONT_POLY=0: k_poly does it)
    bool front_fold = true;  // ... including the FFT probe (ATSC_FRONT_FOLD=0: k_fft_fwd's own probe reads the samples again)
    Engine eng[MAX_ENGINES];
    uint64_t wave_samples = 0;
    double *inv_d2 = nullptr;
    // geometry cache
    std::map<uint32_t, int> geom_idx;
    std::map<uint32_t, uint32_t> pad_cache;
    std::vector<FftGeom> geoms_host;
    std::vector<void *> geom_allocs;
    FftGeom *geoms_dev = nullptr;  // GEOM_CAP entries, appended to (never reallocated: waves in flight read it)
    size_t geoms_uploaded = 0;
    uint64_t launches = 0;
    // multi-device calls: page-locked staging of this device's share of the payload bytes
    uint8_t *h_stage = nullptr;
    size_t h_stage_cap = 0;
    // device span of the last compress / decompress call: first operation issued -> last one done
    cudaEvent_t ev_begin = nullptr, ev_end[MAX_ENGINES] = {};
    double last_call_ms = 0.0;
    // CUDA-event time of every kernel (ms accumulated since reset):
    // 0 stats, 1 plan+poly, 2 rle, 3 fft_fwd, 4 noop+select+scan, 5 emit, 6 decode,
    // 7 HOST time spent preparing and launching waves (not a kernel: shows when a call is host bound),
    // 8 fft_small, 9 fft (top-k + refinement loop), 10 front (fused stats + first polynomial step + probe fold)
    double ms[12] = {};
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

// inside the wave loops: record the error and leave the loop, so that the common tail still drains every
// engine (async copies into the caller's buffers must not outlive the call)
#define CKB(call)                                                                                  \
    {                                                                                              \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            char b__[512];                                                                         \
            snprintf(b__, sizeof b__, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            D.err = b__;                                                                           \
            rc = ATSC_ERR_CUDA;                                                                    \
            break;                                                                                 \
        }                                                                                          \
    }

// grows a buffer that only the engine's own stream touches: the stream is drained first so no
// kernel of an earlier wave still uses the old allocation
template <class T>
int grow(Device &D, cudaStream_t st, T *&p, size_t &cap, size_t need, bool pinned_host = false) {
    if (need <= cap) return ATSC_OK;
    size_t nc = std::max(need, cap + cap / 2);
    if (p) {
        CK(cudaStreamSynchronize(st));
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
    while (n % 9 == 0 && k < sizeof rad) {
        rad[k++] = 9;
        n /= 9;
    }
    while (n % 8 == 0 && k < sizeof rad) {
        rad[k++] = 8;
        n /= 8;
    }
    while (n % 4 == 0 && k < sizeof rad) {
        rad[k++] = 4;
        n /= 4;
    }
    while (n % 2 == 0 && k < sizeof rad) {
        rad[k++] = 2;
        n /= 2;
    }
    while (n % 3 == 0 && k < sizeof rad) {
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
    g.T4T = nullptr;
    if (fast) {
        std::vector<float2> T4((size_t)g.M);
        for (uint32_t n2 = 0; n2 < g.M2; n2++)
            for (uint32_t k1 = 0; k1 < g.M1; k1++) T4[(size_t)n2 * g.M1 + k1] = root((double)(((uint64_t)k1 * n2) % g.M), g.M);
        if ((rc = upload_vec(D, T4, &g.T4))) return rc;
        std::vector<float2> T4T((size_t)g.M);
        for (uint32_t k1 = 0; k1 < g.M1; k1++)
            for (uint32_t n2 = 0; n2 < g.M2; n2++) T4T[(size_t)k1 * g.M2 + n2] = T4[(size_t)n2 * g.M1 + k1];
        if ((rc = upload_vec(D, T4T, &g.T4T))) return rc;
    }
    int idx = (int)D.geoms_host.size();
    D.geoms_host.push_back(g);
    D.geom_idx[L] = idx;
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

// uploads geometries created since the last call (append only)
int sync_geoms(Device &D) {
    if (D.geoms_uploaded == D.geoms_host.size()) return ATSC_OK;
    if (D.geoms_host.size() > GEOM_CAP) {
        D.err = "too many distinct FFT lengths";
        return ATSC_ERR_UNSUPPORTED;
    }
    CK(cudaMemcpyAsync(D.geoms_dev + D.geoms_uploaded, D.geoms_host.data() + D.geoms_uploaded,
                       (D.geoms_host.size() - D.geoms_uploaded) * sizeof(FftGeom), cudaMemcpyHostToDevice, D.st));
    CK(cudaStreamSynchronize(D.st));
    D.geoms_uploaded = D.geoms_host.size();
    return ATSC_OK;
}

// ---------------------------------------------------------------- device setup
void engine_free(Engine &E) {
    SlotPool &P = E.pool;
    void *ptrs[] = {P.rle_k0, P.rle_k1, P.rle_i0, P.rle_i1, P.rle_bnd, P.fft_W, P.fft_Xd, P.fft_keys, P.fft_rank,
                    P.fft_locD, P.fft_locM, P.fft_ovr, P.fft_cD, P.fft_cM, P.fft_dlist, P.poly_slope, P.dec_pts,
                    P.dec_mark, P.dec_idx, E.queues, E.d_ctl, E.d_frames, E.d_samples, E.d_arena, E.d_payload,
                    E.d_dec, E.d_pay_in, E.d_out, E.d_status, E.d_spec_xd, E.d_spec_keys, E.d_chunks, E.d_parts,
                    E.d_items, E.d_fold, E.d_pitems, E.d_ppart, E.d_p1list};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    void *hp[] = {E.h_ctl, E.h_frames, E.h_dec, E.h_status, E.h_chunks, E.h_items, E.h_pitems};
    for (void *p : hp)
        if (p) cudaFreeHost(p);
    for (auto &ev : E.ev)
        if (ev) cudaEventDestroy(ev);
    if (E.st) cudaStreamDestroy(E.st);
    E = Engine{};
}

int engine_alloc(Device &D, Engine &E, int sms) {
    CK(cudaStreamCreateWithFlags(&E.st, cudaStreamNonBlocking));
    SlotPool &P = E.pool;
    P.rle_slots = 2 * sms;
    P.fft_slots = 2 * sms;  // k_fft / k_decode run two 512-thread CTAs per SM
    P.dec_slots = 2 * sms;
    P.fwd_slots = 2 * sms;  // k_fft_fwd: two 288-thread CTAs per SM (96 registers), each with its own W
    size_t rs = (size_t)P.rle_slots;
    CK(cudaMalloc((void **)&P.rle_k0, rs * MAX_FRAME * 8));
    CK(cudaMalloc((void **)&P.rle_k1, rs * MAX_FRAME * 8));
    CK(cudaMalloc((void **)&P.rle_i0, rs * MAX_FRAME * 4));
    CK(cudaMalloc((void **)&P.rle_i1, rs * MAX_FRAME * 4));
    CK(cudaMalloc((void **)&P.rle_bnd, rs * (MAX_FRAME + 8) * 4));
    size_t fs = (size_t)P.fft_slots, hb = MAX_FFT_LEN / 2 + 8;
    CK(cudaMalloc((void **)&P.fft_W, (size_t)std::max(P.fft_slots, P.fwd_slots) * MAX_FFT_LEN * sizeof(float2)));
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
    for (auto &e : E.ev) CK(cudaEventCreate(&e));
    CK(cudaMalloc((void **)&E.queues, 64 * sizeof(unsigned)));
    CK(cudaMalloc((void **)&E.d_ctl, sizeof(WaveCtl)));
    CK(cudaMallocHost((void **)&E.h_ctl, sizeof(WaveCtl)));
    return ATSC_OK;
}

// ~3.4 GB of workspaces per engine.  An allocation can fail in the middle of a call (engines are set up when
// a call first reaches them, and cudaErrorMemoryAllocation is not sticky): the engine is then torn down
// completely, so a later call neither launches on null workspaces nor leaks the part that was allocated.
int engine_init(Device &D, Engine &E, int sms) {
    if (E.ready) return ATSC_OK;
    int rc = engine_alloc(D, E, sms);
    if (rc) {
        cudaGetLastError();
        engine_free(E);
        return rc;
    }
    E.ready = true;
    return ATSC_OK;
}

int env_int(const char *name, int def, int lo, int hi) {
    const char *v = getenv(name);
    if (!v || !*v) return def;
    int x = atoi(v);
    return x < lo ? lo : x > hi ? hi : x;
}

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
    // waves in flight per device and samples per wave (tunable for experiments)
    D.n_engines = env_int("ATSC_ENGINES", 4, 1, MAX_ENGINES);
    D.wave_samples = (uint64_t)env_int("ATSC_WAVE_MI", 72, 1, 512) << 20;
    D.front = env_int("ATSC_FRONT", 2, 0, 2);
    D.front_fold = env_int("ATSC_FRONT_FOLD", 1, 0, 1) != 0;
    D.front_poly = env_int("ATSC_FRONT_POLY", 1, 0, 1) != 0;
    D.poly_items = env_int("ATSC_POLY_ITEMS", 1, 0, 1) != 0;
    D.poly1_static = env_int("ATSC_POLY1_STATIC", 1, 0, 1) != 0;
    D.poly_items_maxf = env_int("ATSC_POLY_ITEMS_MAXF", 3, 0, 1 << 20);
    D.probe_kernel = env_int("ATSC_PROBE_KERNEL", 1, 0, 1) != 0;
    D.sms = sms;
    if ((rc = engine_init(D, D.eng[0], sms))) return rc;  // the others are set up when a call first needs them
    CK(cudaEventCreate(&D.ev_begin));
    for (int e = 0; e < D.n_engines; e++) CK(cudaEventCreate(&D.ev_end[e]));
    CK(cudaMalloc((void **)&D.geoms_dev, GEOM_CAP * sizeof(FftGeom)));
    CK(cudaMalloc((void **)&D.inv_d2, (size_t)(MAX_FRAME + 8) * 8));
    launch_inv_d2(D.inv_d2, MAX_FRAME + 8, D.st);
    D.launches++;
    CK(cudaStreamSynchronize(D.st));
    return ATSC_OK;
}

void device_free(Device &D) {
    cudaSetDevice(D.id);
    cudaDeviceSynchronize();
    for (int e = 0; e < D.n_engines; e++) engine_free(D.eng[e]);
    if (D.ev_begin) cudaEventDestroy(D.ev_begin);
    for (auto &ev : D.ev_end)
        if (ev) cudaEventDestroy(ev);
    if (D.h_stage) cudaFreeHost(D.h_stage);
    if (D.inv_d2) cudaFree(D.inv_d2);
    if (D.geoms_dev) cudaFree(D.geoms_dev);
    for (void *p : D.geom_allocs) cudaFree(p);
    if (D.st) cudaStreamDestroy(D.st);
}

// device memory (or managed): *device = the GPU that owns it
bool is_device_ptr(const void *p, int *device = nullptr) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (device) *device = a.device;
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
// the context's Device that owns a device-resident buffer; nullptr when that GPU is not part of the context
Device *owner_of(atsc_ctx *ctx, int device);

// device span of a call: begin is recorded on the first engine's stream before anything is issued,
// an end event on every engine's stream after its last operation
int span_begin(Device &D) {
    CK(cudaEventRecord(D.ev_begin, D.eng[0].st));
    return ATSC_OK;
}
int span_end(Device &D) {
    double ms = 0.0;
    for (int k = 0; k < D.n_engines; k++)
        if (D.eng[k].st) CK(cudaEventRecord(D.ev_end[k], D.eng[k].st));
    for (int k = 0; k < D.n_engines; k++) {
        if (!D.eng[k].st) continue;  // engine never reached by a call yet
        CK(cudaStreamSynchronize(D.eng[k].st));
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, D.ev_begin, D.ev_end[k]));
        ms = std::max(ms, (double)t);
    }
    D.last_call_ms = ms;
    return ATSC_OK;
}

// ---------------------------------------------------------------- compress
// Issues the whole pipeline of one wave on E's stream without waiting for it.  The frames'
// samples live at d_samples (device).  collect_wave() picks the results up.
struct HostTimer {
    double &acc;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    explicit HostTimer(double &a) : acc(a) {}
    ~HostTimer() { acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

int issue_wave(Device &D, Engine &E, const double *d_samples, const std::vector<FrameReq> &reqs,
               float max_error_f32) {
    HostTimer host_timer(D.ms[7]);
    const uint32_t n = (uint32_t)reqs.size();
    const double max_err = (double)max_error_f32;  // `max_error as f64` (frame/mod.rs:67,87)
    int rc;
    size_t hcap = E.frames_cap;
    if ((rc = grow(D, E.st, E.d_frames, E.frames_cap, n))) return rc;
    if ((rc = grow(D, E.st, E.h_frames, hcap, E.frames_cap, true))) return rc;
    uint64_t arena = 0, spec = 0, samples = 0;
    bool any_noop = false, any_small = false, any_large = false;
    uint32_t small_lmax = 2;
    // frames a front-end kernel takes: long enough, and 16-byte aligned pairs (bulk copies, double2 reads)
    const bool base_ok = D.front && ((uintptr_t)d_samples & 15u) == 0;
    // frames whose FFT probe can be folded while they stream: bounded Auto, padded length 2^a * 3^7 with an even
    // number of replicated pairs in front (fft2.cuh, fold_acc)
    auto fold_frame = [&](const FrameReq &r) {
        if (!(r.bounded && r.comp == C_AUTO && r.forced == 0xFF && !r.select_only && r.len >= 128)) return false;
        const uint32_t L = padded_len(D, r.len);
        return L % (2u * 243u) == 0 && f2_fold_ra(L / (2u * 243u)) != 0 && ((L - r.len) / 2) % 2 == 0;
    };
    auto front_frame = [&](const FrameReq &r) {
        if (!(base_ok && r.len >= FRONT_MIN_SAMPLES && (r.len & 1u) == 0 && (r.off & 1ull) == 0)) return false;
        if (D.front == 2) return fold_frame(r);  // k_sfold
        // (the first Polynomial step of such a frame is 100 samples: N / max(3, N / 100), polynomial.rs:218-221)
        return r.len / std::max<uint32_t>(3, r.len / 100) == 100;
    };
    // k_stats chunks sit in [0, n_chunks) of the chunk table, k_sfold's slot-range items behind them; every entry
    // has one StatsPart
    size_t n_chunks = 0, n_items = 0, n_sf = 0, n_sf_frames = 0;
    std::vector<uint8_t> &fronted = E.fronted;  // per frame: does a front-end kernel take it
    fronted.assign(n, 0);
    for (uint32_t i = 0; i < n; i++) {
        if (!D.front || !front_frame(reqs[i])) {
            n_chunks += (reqs[i].len + STATS_CHUNK - 1) / STATS_CHUNK;
            continue;
        }
        fronted[i] = 1;
        if (D.front == 2) {
            n_sf += sfold_items(padded_len(D, reqs[i].len) / (2u * 243u));
            n_sf_frames++;
        } else {
            n_items++;
        }
    }
    // frames whose first Polynomial step k_poly1s (or k_poly1) evaluates in balanced work items (poly.cuh)
    auto poly1_frame = [&](const FrameReq &r) {
        return D.poly_items && r.len >= POLY_ITEM_MIN_LEN && poly_first_step(r.len) == P1_STEP && r.bounded && !r.select_only &&
               (r.comp == C_POLY || (r.comp == C_AUTO && (r.forced == 0xFF || r.forced == C_POLY)));
    };
    // (only when big frames are scarce: with several frames per CTA slot whole-frame items balance by themselves
    // and the extra launch costs more than it saves -- 3072 x 64 k Polynomial fleet: 2.17 vs 1.90 ms)
    size_t n_pitems = 0, n_pframes = 0;
    if (D.front != 1)
        for (uint32_t i = 0; i < n; i++)
            if (poly1_frame(reqs[i])) {
                n_pitems += poly_item_count(reqs[i].len);
                n_pframes++;
            }
    if (n_pframes > (size_t)D.poly_items_maxf * (size_t)E.pool.poly_slots) n_pitems = 0;
    hcap = E.pitems_cap;
    if ((rc = grow(D, E.st, E.d_pitems, E.pitems_cap, n_pitems))) return rc;
    if ((rc = grow(D, E.st, E.h_pitems, hcap, E.pitems_cap, true))) return rc;
    if ((rc = grow(D, E.st, E.d_ppart, E.ppart_cap, n_pitems * (D.poly1_static ? (size_t)P1_PARTS : 1u)))) return rc;
    if (D.poly1_static && (rc = grow(D, E.st, E.d_p1list, E.p1list_cap, n_pitems))) return rc;
    hcap = E.chunks_cap;
    if ((rc = grow(D, E.st, E.d_chunks, E.chunks_cap, n_chunks + n_sf))) return rc;
    if ((rc = grow(D, E.st, E.h_chunks, hcap, E.chunks_cap, true))) return rc;
    if ((rc = grow(D, E.st, E.d_parts, E.parts_cap, n_chunks + n_sf))) return rc;
    if ((rc = grow(D, E.st, E.d_fold, E.fold_cap, n_sf_frames * (size_t)SF_FOLD_SLOTS))) return rc;
    hcap = E.items_cap;
    if (D.front == 2) n_items = n_sf_frames;  // k_probe's list: the frames k_sfold folds
    if ((rc = grow(D, E.st, E.d_items, E.items_cap, n_items))) return rc;
    if ((rc = grow(D, E.st, E.h_items, hcap, E.items_cap, true))) return rc;
    uint32_t nc = 0, ni = 0, nsf = 0, nsf_frames = 0, npi = 0;
    for (uint32_t i = 0; i < n; i++) {
        FrameWork &f = E.h_frames[i];
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
        f.chunk0 = nc;
        if (D.front == 2 && fronted[i]) {
            f.front_mode = FM_SFOLD;
            f.chunk0 = (uint32_t)n_chunks + nsf;
            f.fold_slot = nsf_frames++;
            E.h_items[ni++] = i;
            const uint32_t slots = sfold_slots(padded_len(D, r.len) / (2u * 243u));
            for (uint32_t s0 = 0; s0 < slots; s0 += SF_ITEM) E.h_chunks[n_chunks + nsf++] = ChunkRef{i, s0};
        } else if (fronted[i]) {
            f.front_mode = FM_ON;
            // the first Polynomial step is worth evaluating when the bounded Catmull-Rom loop will run
            if (D.front == 1 && D.front_poly && r.bounded && !r.select_only &&
                (r.comp == C_POLY || (r.comp == C_AUTO && (r.forced == 0xFF || r.forced == C_POLY))))
                f.front_mode |= FM_POLY;
            E.h_items[ni++] = i;
        } else {
            for (uint32_t c0 = 0; c0 < r.len; c0 += STATS_CHUNK) E.h_chunks[nc++] = ChunkRef{i, c0};
        }
        if (n_pitems && poly1_frame(r)) {
            f.poly_part0 = npi;
            f.poly_parts = poly_item_count(r.len);
            for (uint32_t q = 0; q < f.poly_parts; q++) E.h_pitems[npi++] = ChunkRef{i, q};
        }
        samples += r.len;
        uint8_t eff = (r.comp == C_AUTO && r.forced != 0xFF) ? r.forced : r.comp;
        any_noop |= r.comp == C_NOOP;
        if (eff == C_FFT || eff == C_AUTO) {
            uint32_t L = (r.bounded && r.len >= 128) ? padded_len(D, r.len) : r.len;
            f.fft_small = L <= 1152;
            any_small |= L <= 1152;
            if (L <= 1152) small_lmax = std::max(small_lmax, L);
            any_large |= L > 1152;
            if (r.len >= 128) {
                int gi;
                if ((rc = get_geom(D, L, &gi))) {
                    D.err = "FFT length is not of the form 2^a*3^b (unbounded Compressor::FFT needs a decomposable frame length)";
                    return rc;
                }
                f.geom = gi;
                const FftGeom &G = D.geoms_host[gi];
                if (G.T4) {
                    f.spec_off = spec;
                    spec += G.M + 8;
                    // Auto frames that k_front streams also get the FFT probe there
                    if (D.front_fold && (f.front_mode & FM_ON) && r.bounded && r.comp == C_AUTO && r.forced == 0xFF && !r.select_only &&
                        f2_fold_ra(G.M1) && ((L - r.len) / 2) % 2 == 0) {
                        f.front_mode |= FM_FOLD;
                    }
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
    if ((rc = grow(D, E.st, E.d_arena, E.arena_cap, (size_t)arena + 1))) return rc;
    if ((rc = grow(D, E.st, E.d_spec_xd, E.spec_xd_cap, (size_t)spec + 1))) return rc;
    if ((rc = grow(D, E.st, E.d_spec_keys, E.spec_keys_cap, (size_t)spec + 1))) return rc;
    // payload buffer sized before the total is known: one byte per sample covers every fleet but
    // near-incompressible ones, which take the overflow path of collect_wave()
    if ((rc = grow(D, E.st, E.d_payload, E.payload_cap, (size_t)std::max<uint64_t>(samples, 1u << 20) + 64 * (size_t)n))) return rc;
    cudaStream_t st = E.st;
    CK(cudaMemcpyAsync(E.d_frames, E.h_frames, (size_t)n * sizeof(FrameWork), cudaMemcpyHostToDevice, st));
    if (n_chunks + n_sf) CK(cudaMemcpyAsync(E.d_chunks, E.h_chunks, (n_chunks + n_sf) * sizeof(ChunkRef), cudaMemcpyHostToDevice, st));
    if (n_items) CK(cudaMemcpyAsync(E.d_items, E.h_items, n_items * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    if (n_pitems && !D.poly1_static) CK(cudaMemcpyAsync(E.d_pitems, E.h_pitems, n_pitems * sizeof(ChunkRef), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(E.queues, 0, 64 * sizeof(unsigned), st));
    CK(cudaMemsetAsync(E.d_ctl, 0, sizeof(WaveCtl), st));
    CK(cudaEventRecord(E.ev[12], st));
    if (n_items && D.front == 1) {
        launch_front(E.d_frames, E.d_items, (uint32_t)n_items, d_samples, max_err, D.geoms_dev, E.pool, E.queues + 9, st);
        D.launches++;
    }
    if (n_sf) {
        launch_sfold(E.d_frames, E.d_chunks + n_chunks, (uint32_t)n_sf, d_samples, D.geoms_dev, E.d_fold, E.d_parts, E.queues + 9, st);
        D.launches++;
    }
    CK(cudaEventRecord(E.ev[13], st));
    CK(cudaEventRecord(E.ev[0], st));
    if (n_chunks) {
        launch_stats(E.d_frames, E.d_chunks, (uint32_t)n_chunks, d_samples, E.d_parts, E.queues + 0, st);
        D.launches++;
    }
    CK(cudaEventRecord(E.ev[1], st));
    const bool p1s = n_pitems && D.poly1_static;
    launch_plan(E.d_frames, n, d_samples, E.d_parts, D.geoms_dev, p1s ? E.d_p1list : nullptr, E.queues + 11, st);
    if (n_pitems) {
        if (p1s) {
            launch_poly1s(E.d_p1list, E.queues + 11, (uint32_t)n_pitems, E.d_ppart, st);
            D.launches++;
        } else {
            launch_poly1(E.d_frames, E.d_pitems, (uint32_t)n_pitems, d_samples, E.d_ppart, E.queues + 11, st);
            D.launches++;
        }
    }
    launch_poly(E.d_frames, n, d_samples, max_err, D.inv_d2, E.pool, E.d_ppart, D.poly1_static ? P1_PARTS : 1u, E.queues + 1, st);
    CK(cudaEventRecord(E.ev[2], st));
    if (any_small) {
        launch_fft_small(E.d_frames, n, d_samples, max_err, D.geoms_dev, E.d_arena, small_lmax, E.queues + 8, st);
        D.launches++;
    }
    CK(cudaEventRecord(E.ev[10], st));
    if (spec && n_sf && D.probe_kernel) {
        launch_probe(E.d_frames, E.d_items, (uint32_t)n_items, max_err, D.geoms_dev, E.d_fold, E.queues + 12, st);
        D.launches++;
    }
    if (spec) {
        launch_fft_fwd(E.d_frames, n, d_samples, max_err, D.geoms_dev, E.pool, E.d_spec_xd, E.d_spec_keys, E.d_fold, E.queues + 7, st);
        D.launches++;
    }
    CK(cudaEventRecord(E.ev[11], st));
    if (any_large) {
        launch_fft(E.d_frames, n, d_samples, max_err, D.geoms_dev, E.pool, E.d_arena, E.d_spec_xd, E.d_spec_keys, E.queues + 3, st);
        D.launches++;
    }
    CK(cudaEventRecord(E.ev[3], st));
    // RLE last: its sort runs only for frames where a size lower bound still beats Polynomial and FFT
    launch_rle(E.d_frames, n, d_samples, max_err, E.pool, E.queues + 2, st);
    CK(cudaEventRecord(E.ev[4], st));
    D.launches += 3;
    if (any_noop) {
        launch_noop_size(E.d_frames, n, d_samples, E.queues + 4, st);
        D.launches++;
    }
    launch_select(E.d_frames, n, max_err, st);
    launch_scan(E.d_frames, n, &E.d_ctl->total, st);
    D.launches += 2;
    CK(cudaEventRecord(E.ev[5], st));
    CK(cudaEventRecord(E.ev[6], st));
    launch_emit(E.d_frames, n, d_samples, D.geoms_dev, E.pool, E.d_arena, E.d_payload, &E.d_ctl->total,
                (unsigned long long)E.payload_cap, &E.d_ctl->overflow, E.queues + 5, st);
    CK(cudaEventRecord(E.ev[7], st));
    D.launches++;
    CK(cudaMemcpyAsync(E.h_frames, E.d_frames, (size_t)n * sizeof(FrameWork), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(E.h_ctl, E.d_ctl, sizeof(WaveCtl), cudaMemcpyDeviceToHost, st));
    CK(cudaGetLastError());
    return ATSC_OK;
}

// Waits for the wave issued on E; afterwards E.h_frames[0..n) hold the results and the wave's
// payload bytes sit in E.d_payload[0..*payload_total) (the overflow path re-emits into a larger buffer).
int wait_wave(Device &D, Engine &E, uint32_t n, const double *d_samples, uint64_t *payload_total) {
    CK(cudaStreamSynchronize(E.st));
    CK(cudaGetLastError());
    {
        // stream order: stats | plan+poly | fft_small | fft_fwd | fft | rle | noop+select+scan
        const int a[7] = {0, 1, 2, 10, 11, 3, 4}, b[7] = {1, 2, 10, 11, 3, 4, 5}, slot[7] = {0, 1, 8, 3, 9, 2, 4};
        for (int k = 0; k < 7; k++) {
            float t = 0.f;
            CK(cudaEventElapsedTime(&t, E.ev[a[k]], E.ev[b[k]]));
            D.ms[slot[k]] += t;
        }
    }
    float te = 0.f;
    CK(cudaEventElapsedTime(&te, E.ev[6], E.ev[7]));
    D.ms[5] += te;
    CK(cudaEventElapsedTime(&te, E.ev[12], E.ev[13]));
    D.ms[10] += te;
    const uint64_t total = E.h_ctl->total;
    *payload_total = total;
    if (E.h_ctl->overflow) {
        int rc = grow(D, E.st, E.d_payload, E.payload_cap, (size_t)total + 64);
        if (rc) return rc;
        CK(cudaMemsetAsync(E.queues + 5, 0, sizeof(unsigned), E.st));
        CK(cudaMemsetAsync(&E.d_ctl->overflow, 0, sizeof(unsigned), E.st));
        launch_emit(E.d_frames, n, d_samples, D.geoms_dev, E.pool, E.d_arena, E.d_payload, &E.d_ctl->total,
                    (unsigned long long)E.payload_cap, &E.d_ctl->overflow, E.queues + 5, E.st);
        D.launches++;
        CK(cudaStreamSynchronize(E.st));
        CK(cudaGetLastError());
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
};

static const uint32_t COMPRESSION_SPEED[7] = {2147483647u, 4096, 2048, 1024, 512, 256, 128};  // frame/mod.rs:22

// collects the wave pending on E: frame records -> out[], payload bytes -> sink (in wave order)
int collect_wave(Device &D, Engine &E, const uint32_t *idx, atsc_frame_out *out, PayloadSink &sink) {
    WaveJob &J = E.job;
    if (!J.active) return ATSC_OK;
    J.active = false;
    uint64_t ptotal = 0;
    int rc = wait_wave(D, E, J.n, J.d_samples, &ptotal);
    if (rc) return rc;
    for (uint32_t k = 0; k < J.n; k++) {
        const FrameWork &f = E.h_frames[k];
        atsc_frame_out &o = out[idx[J.pos + k]];
        fill_out(f, o, sink.used);
        if (J.sampled && J.reqs[k].forced != 0xFF && !f.is_const) {
            o.near_tie |= J.tie1[k];
            for (int c = 0; c < 3; c++) {
                o.cand_error[c] = J.diag1[k].cand_error[c];
                o.cand_size[c] = J.diag1[k].cand_size[c];
            }
        }
    }
    if (sink.used + ptotal > sink.cap)
        sink.overflow = true;
    else if (ptotal)
        // page-locked destination: asynchronous, overlapped with the other engines' waves (the
        // caller drains every stream before returning); pageable: staged by the runtime
        CK(cudaMemcpyAsync(sink.buf + sink.used, E.d_payload, ptotal, cudaMemcpyDeviceToHost, E.st));
    sink.used += ptotal;
    return ATSC_OK;
}

// compress frames idx[0..m) on device D
int compress_on_device(Device &D, const double *samples, bool dev_ptr, const uint64_t *frame_off,
                       const uint32_t *frame_len, const uint32_t *idx, uint32_t m, uint8_t compressor,
                       float max_error, uint32_t speed, int bounded, atsc_frame_out *out, PayloadSink &sink) {
    CK(cudaSetDevice(D.id));
    const bool sampled = bounded && compressor == C_AUTO && speed > 0;
    const uint32_t sample_n = COMPRESSION_SPEED[speed];
    uint32_t pos = 0, wave = 0;
    std::vector<FrameReq> sel;
    std::vector<uint32_t> sel_of;
    int rc = span_begin(D);
    if (rc) return rc;
    while (pos < m) {
        // ---- cut a wave
        uint64_t tot = 0;
        uint32_t end = pos;
        uint64_t lo = ~0ull, hi = 0;
        while (end < m && end - pos < WAVE_FRAMES) {
            uint32_t fi = idx[end];
            if (tot && tot + frame_len[fi] > D.wave_samples) break;
            tot += frame_len[fi];
            lo = std::min(lo, frame_off[fi]);
            hi = std::max(hi, frame_off[fi] + frame_len[fi]);
            end++;
        }
        const uint32_t n = end - pos;
        Engine &E = D.eng[wave % D.n_engines];
        wave++;
        if ((rc = engine_init(D, E, D.sms))) break;  // ~3.4 GB of workspaces, only for engines a call reaches
        // the engine's previous wave must be collected before its buffers are reused; waves are
        // collected in issue order, so payloads land in frame order
        if ((rc = collect_wave(D, E, idx, out, sink))) break;
        // ---- samples on the device
        const double *d_samples = samples;
        bool packed = false;
        if (!dev_ptr) {
            uint64_t span = hi - lo;
            packed = span > 2 * tot + 4096;
            if ((rc = grow(D, E.st, E.d_samples, E.samples_cap, (size_t)(packed ? tot : span) + 8))) break;
            if (!packed) {
                CKB(cudaMemcpyAsync(E.d_samples, samples + lo, span * 8, cudaMemcpyHostToDevice, E.st));
            } else {
                uint64_t o = 0;
                for (uint32_t k = pos; k < end && !rc; k++) {
                    uint32_t fi = idx[k];
                    cudaError_t ce = cudaMemcpyAsync(E.d_samples + o, samples + frame_off[fi], (size_t)frame_len[fi] * 8,
                                                     cudaMemcpyHostToDevice, E.st);
                    if (ce != cudaSuccess) {
                        D.err = std::string("cudaMemcpyAsync (packed samples): ") + cudaGetErrorString(ce);
                        rc = ATSC_ERR_CUDA;
                    }
                    o += frame_len[fi];
                }
                if (rc) break;
            }
            d_samples = E.d_samples;
        }
        WaveJob &J = E.job;
        J.reqs.clear();
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
            J.reqs.push_back(r);
        }
        J.sampled = sampled;
        J.tie1.assign(n, 0);
        J.diag1.clear();
        if (sampled) {
            // frame/mod.rs:89-111: pick the compressor on data[0..sample], then compress everything with it
            sel.clear();
            sel_of.clear();
            for (uint32_t k = 0; k < n; k++) {
                if (J.reqs[k].len >= sample_n) {
                    FrameReq r = J.reqs[k];
                    r.len = sample_n;
                    r.select_only = 1;
                    sel.push_back(r);
                    sel_of.push_back(k);
                }
            }
            if (!sel.empty()) {
                uint64_t pt;
                if ((rc = issue_wave(D, E, d_samples, sel, max_error))) break;
                if ((rc = wait_wave(D, E, (uint32_t)sel.size(), d_samples, &pt))) break;
                J.diag1.resize(n);
                for (size_t s = 0; s < sel.size(); s++) {
                    const FrameWork &f = E.h_frames[s];
                    J.reqs[sel_of[s]].forced = f.winner;
                    J.tie1[sel_of[s]] = f.near_tie;
                    fill_out(f, J.diag1[sel_of[s]], 0);
                }
            } else {
                J.diag1.resize(n);
            }
        }
        if ((rc = issue_wave(D, E, d_samples, J.reqs, max_error))) break;
        J.active = true;
        J.pos = pos;
        J.n = n;
        J.d_samples = d_samples;
        pos = end;
    }
    // ---- drain: collect the remaining waves in issue order, then wait for the payload copies
    for (int k = 0; k < D.n_engines; k++) {
        Engine &E = D.eng[(wave + k) % D.n_engines];
        int rc2 = rc ? ATSC_OK : collect_wave(D, E, idx, out, sink);
        if (rc2) rc = rc2;
        E.job.active = false;
    }
    int rc3 = span_end(D);  // synchronises every engine's stream, also on the error path
    if (!rc) rc = rc3;
    if (rc)
        for (int k = 0; k < D.n_engines; k++)
            if (D.eng[k].st) cudaStreamSynchronize(D.eng[k].st);
    if (!rc) CK(cudaGetLastError());
    return rc;
}

// ---------------------------------------------------------------- decompress
// waits for the decode wave pending on E and checks its per-frame status words
int collect_decode(Device &D, Engine &E, const uint32_t *idx) {
    if (!E.dec_active) return ATSC_OK;
    E.dec_active = false;
    CK(cudaStreamSynchronize(E.st));
    CK(cudaGetLastError());
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, E.ev[8], E.ev[9]));
    D.ms[6] += t;
    for (uint32_t k = 0; k < E.dec_n; k++)
        if (E.h_status[k]) {
            char b[128];
            snprintf(b, sizeof b, "frame %u: malformed or unsupported payload (code %u)", idx[E.dec_pos + k], E.h_status[k]);
            D.err = b;
            return E.h_status[k] == 4 ? ATSC_ERR_UNSUPPORTED : ATSC_ERR_FORMAT;
        }
    return ATSC_OK;
}

// Waves of frames go round the engines like the compress path: payload H2D, k_decode and the
// output D2H of one wave overlap the other engines' waves.
int decompress_on_device(Device &D, const atsc_frame_in *frames, const uint32_t *idx, uint32_t m,
                         const uint8_t *payloads, uint64_t payload_bytes, double *out, bool out_dev) {
    CK(cudaSetDevice(D.id));
    uint32_t pos = 0, wave = 0;
    int rc = span_begin(D);
    while (pos < m && !rc) {
        uint64_t tot = 0;
        uint32_t end = pos;
        uint64_t plo = ~0ull, phi = 0;
        while (end < m && end - pos < WAVE_FRAMES) {
            const atsc_frame_in &f = frames[idx[end]];
            if (tot && tot + f.sample_count > D.wave_samples) break;
            tot += f.sample_count;
            plo = std::min<uint64_t>(plo, f.payload_off);
            phi = std::max<uint64_t>(phi, f.payload_off + f.payload_len);
            end++;
        }
        const uint32_t n = end - pos;
        if (phi > payload_bytes) {
            D.err = "frame payload range exceeds payload_bytes";
            rc = ATSC_ERR_ARG;
            break;
        }
        Engine &E = D.eng[wave % D.n_engines];
        wave++;
        if ((rc = engine_init(D, E, D.sms))) break;
        if ((rc = collect_decode(D, E, idx))) break;
        size_t hc = E.dec_cap;
        if ((rc = grow(D, E.st, E.d_dec, E.dec_cap, n))) break;
        if ((rc = grow(D, E.st, E.h_dec, hc, E.dec_cap, true))) break;
        hc = E.status_cap;
        if ((rc = grow(D, E.st, E.d_status, E.status_cap, n))) break;
        if ((rc = grow(D, E.st, E.h_status, hc, E.status_cap, true))) break;
        if ((rc = grow(D, E.st, E.d_pay_in, E.pay_in_cap, (size_t)(phi - plo) + 64))) break;
        if (!out_dev && (rc = grow(D, E.st, E.d_out, E.out_cap, (size_t)tot + 8))) break;
        uint64_t oo = 0;
        for (uint32_t k = 0; k < n && !rc; k++) {
            const atsc_frame_in &f = frames[idx[pos + k]];
            DecFrame &d = E.h_dec[k];
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
                rc = ATSC_ERR_ARG;
            } else if (f.compressor == C_FFT && f.sample_count >= 128) {
                int gi;
                if (!(rc = get_geom(D, padded_len(D, f.sample_count), &gi))) d.geom = gi;
            }
        }
        if (rc || (rc = sync_geoms(D))) break;
        CKB(cudaMemcpyAsync(E.d_dec, E.h_dec, (size_t)n * sizeof(DecFrame), cudaMemcpyHostToDevice, E.st));
        CKB(cudaMemcpyAsync(E.d_pay_in, payloads + plo, (size_t)(phi - plo), cudaMemcpyHostToDevice, E.st));
        CKB(cudaMemsetAsync(E.queues, 0, 64 * sizeof(unsigned), E.st));
        double *d_out = out_dev ? out : E.d_out;
        CKB(cudaEventRecord(E.ev[8], E.st));
        launch_decode(E.d_dec, n, E.d_pay_in, d_out, D.geoms_dev, E.pool, D.inv_d2, E.d_status, E.queues + 6, E.st);
        CKB(cudaEventRecord(E.ev[9], E.st));
        D.launches++;
        CKB(cudaMemcpyAsync(E.h_status, E.d_status, (size_t)n * 4, cudaMemcpyDeviceToHost, E.st));
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
                cudaError_t ce = cudaMemcpyAsync(out + dst, E.d_out + src, len * 8, cudaMemcpyDeviceToHost, E.st);
                if (ce != cudaSuccess) {
                    D.err = std::string("cudaMemcpyAsync (decoded samples): ") + cudaGetErrorString(ce);
                    rc = ATSC_ERR_CUDA;
                    break;
                }
                src += len;
                k = j;
            }
            if (rc) break;
        }
        E.dec_active = true;
        E.dec_pos = pos;
        E.dec_n = n;
        pos = end;
    }
    for (int k = 0; k < D.n_engines; k++) {
        Engine &E = D.eng[(wave + k) % D.n_engines];
        int rc2 = rc ? ATSC_OK : collect_decode(D, E, idx);
        if (rc2) rc = rc2;
        E.dec_active = false;
    }
    int rc3 = span_end(D);  // synchronises every engine's stream, also on the error path
    if (!rc) rc = rc3;
    if (rc)
        for (int k = 0; k < D.n_engines; k++)
            if (D.eng[k].st) cudaStreamSynchronize(D.eng[k].st);
    return rc;
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

Device *owner_of(atsc_ctx *ctx, int device) {
    for (Device *D : ctx->devs)
        if (D->id == device) return D;
    return nullptr;
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

double atsc_gpu_last_call_ms(const atsc_ctx *ctx) {
    double ms = 0.0;
    if (ctx)
        for (Device *D : ctx->devs) ms = std::max(ms, D->last_call_ms);
    return ms;
}

void atsc_gpu_kernel_ms(atsc_ctx *ctx, double *out12, int reset) {
    double *out8 = out12;
    for (int k = 0; k < 12; k++) out8[k] = 0.0;
    if (!ctx) return;
    for (Device *D : ctx->devs)
        for (int k = 0; k < 12; k++) {
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
    int owner = -1;
    const bool dev_ptr = is_device_ptr(samples, &owner);
    const size_t nd = dev_ptr ? 1 : ctx->devs.size();
    if (nd == 1) {
        // device-resident samples are compressed by the GPU that owns them
        Device *Dp = dev_ptr ? owner_of(ctx, owner) : ctx->devs[0];
        if (!Dp) {
            ctx->err = "samples live on a GPU that is not part of this context";
            return ATSC_ERR_ARG;
        }
        std::vector<uint32_t> idx(n_frames);
        for (uint32_t i = 0; i < n_frames; i++) idx[i] = i;
        PayloadSink sink{payload_buf, payload_cap, 0, false};
        Device &D = *Dp;
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
    // Several devices: contiguous frame ranges, one host thread per device, no collective (frames never
    // communicate; the reference stitches sequentially, data.rs:73-75).  Each device's payload lands in its
    // own page-locked staging buffer (~2 B per sample, grown and the share redone if a fleet needs more),
    // then the shares are laid end to end in the caller's buffer.
    auto parts = shard(frame_len, n_frames, nd);
    std::vector<PayloadSink> sinks(nd);
    std::vector<int> rcs(nd, 0);
    std::vector<std::thread> th;
    for (size_t d = 0; d < nd; d++) {
        th.emplace_back([&, d]() {
            if (parts[d].empty()) return;
            Device &D = *ctx->devs[d];
            uint64_t ns = 0;
            for (uint32_t i : parts[d]) ns += frame_len[i];
            uint64_t cap = ns * 2 + 64 * (uint64_t)parts[d].size() + (1u << 20);
            for (int attempt = 0; attempt < 2; attempt++) {
                if (cap > D.h_stage_cap) {
                    cudaSetDevice(D.id);
                    if (D.h_stage) cudaFreeHost(D.h_stage);
                    D.h_stage = nullptr;
                    D.h_stage_cap = 0;
                    if (cudaMallocHost((void **)&D.h_stage, cap) != cudaSuccess) {
                        cudaGetLastError();
                        D.err = "cannot allocate page-locked payload staging";
                        rcs[d] = ATSC_ERR_CUDA;
                        return;
                    }
                    D.h_stage_cap = cap;
                }
                sinks[d] = PayloadSink{D.h_stage, D.h_stage_cap, 0, false};
                rcs[d] = compress_on_device(D, samples, false, frame_off, frame_len, parts[d].data(), (uint32_t)parts[d].size(),
                                            compressor, max_error, speed, bounded, out, sinks[d]);
                if (rcs[d] || !sinks[d].overflow) return;
                cap = sinks[d].used + 64;  // the share needs this much: once more
            }
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
            memcpy(payload_buf + used, sinks[d].buf, sinks[d].used);
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
    int owner = -1;
    const bool out_dev = is_device_ptr(out_samples, &owner);
    const size_t nd = out_dev ? 1 : ctx->devs.size();
    Device *D0 = out_dev ? owner_of(ctx, owner) : ctx->devs[0];
    if (!D0) {
        ctx->err = "out_samples lives on a GPU that is not part of this context";
        return ATSC_ERR_ARG;
    }
    std::vector<uint32_t> lens(n_frames);
    for (uint32_t i = 0; i < n_frames; i++) lens[i] = frames[i].sample_count;
    auto parts = shard(lens.data(), n_frames, nd);
    std::vector<int> rcs(nd, 0);
    if (nd == 1) {
        rcs[0] = decompress_on_device(*D0, frames, parts[0].data(), n_frames, payloads, payload_bytes, out_samples, out_dev);
        if (rcs[0]) {
            ctx->err = D0->err;
            return rcs[0];
        }
        return ATSC_OK;
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

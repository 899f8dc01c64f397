/*
 * atsc_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the hot path of instaclustr/atsc v0.7.2 (Rust):
 * per-frame compressor selection / fitting and decompression, plus the BRO
 * stream layout.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (libatsc_gpu.so) never links, imports or calls it.
 *
 * Parity pinning: every golden vector of the reference's own unit tests for
 * this path plus the demo HTML arrays are checked in
 * tests/test_oracle_golden.py.  Third-party crates that carry arithmetic and
 * are absent from /root/reference are restated from their published
 * algorithms:
 *   rustfft 6.2.0                -> unnormalised complex DFT in f32 (any correct
 *                                   f32 DFT; butterfly order is build-dependent
 *                                   in rustfft itself)
 *   splines 4.3.1                -> Key / Linear / CatmullRom (cubic Hermite)
 *   inverse_distance_weight 0.1.1-> power-2 IDW over all points
 *   bincode 2.0.0-rc.3 standard  -> little-endian + varint
 *   std::collections::BinaryHeap -> rebuild / pop(sift_down_to_bottom+sift_up)
 *
 * Build: gcc -O2 -ffp-contract=off (NO fma contraction: rustc never fuses).
 *
 * Each function cites the reference file:line (relative to
 * /root/reference/atsc/src/) it follows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* byte buffer                                                        */
/* ------------------------------------------------------------------ */
typedef struct {
    uint8_t *p;
    size_t len, cap;
} buf_t;

static void buf_reserve(buf_t *b, size_t extra) {
    if (b->len + extra <= b->cap) return;
    size_t nc = b->cap ? b->cap * 2 : 256;
    while (nc < b->len + extra) nc *= 2;
    b->p = (uint8_t *)realloc(b->p, nc);
    b->cap = nc;
}
static void buf_put(buf_t *b, const void *src, size_t n) {
    buf_reserve(b, n);
    memcpy(b->p + b->len, src, n);
    b->len += n;
}
static void buf_u8(buf_t *b, uint8_t v) { buf_put(b, &v, 1); }
static void buf_free(buf_t *b) {
    free(b->p);
    b->p = NULL;
    b->len = b->cap = 0;
}

/* bincode 2.0.0-rc.3 config::standard(): varint, little endian
 * (compressor/mod.rs:126-130). */
static void enc_varint(buf_t *b, uint64_t u) {
    if (u < 251) {
        buf_u8(b, (uint8_t)u);
    } else if (u < 65536ull) {
        buf_u8(b, 251);
        uint16_t v = (uint16_t)u;
        buf_put(b, &v, 2);
    } else if (u < 4294967296ull) {
        buf_u8(b, 252);
        uint32_t v = (uint32_t)u;
        buf_put(b, &v, 4);
    } else {
        buf_u8(b, 253);
        buf_put(b, &u, 8);
    }
}
static void enc_zigzag(buf_t *b, int64_t n) {
    uint64_t u = ((uint64_t)n << 1) ^ (uint64_t)(n >> 63);
    enc_varint(b, u);
}
static void enc_f32(buf_t *b, float v) { buf_put(b, &v, 4); }
static void enc_f64(buf_t *b, double v) { buf_put(b, &v, 8); }

typedef struct {
    const uint8_t *p;
    size_t len, pos;
    int err;
} rd_t;

static uint8_t rd_u8(rd_t *r) {
    if (r->pos + 1 > r->len) {
        r->err = 1;
        return 0;
    }
    return r->p[r->pos++];
}
static void rd_raw(rd_t *r, void *dst, size_t n) {
    if (r->pos + n > r->len) {
        r->err = 1;
        memset(dst, 0, n);
        return;
    }
    memcpy(dst, r->p + r->pos, n);
    r->pos += n;
}
static uint64_t rd_varint(rd_t *r) {
    uint8_t t = rd_u8(r);
    if (t < 251) return t;
    if (t == 251) {
        uint16_t v;
        rd_raw(r, &v, 2);
        return v;
    }
    if (t == 252) {
        uint32_t v;
        rd_raw(r, &v, 4);
        return v;
    }
    if (t == 253) {
        uint64_t v;
        rd_raw(r, &v, 8);
        return v;
    }
    r->err = 1;
    return 0;
}
static int64_t rd_zigzag(rd_t *r) {
    uint64_t u = rd_varint(r);
    return (int64_t)(u >> 1) ^ -(int64_t)(u & 1);
}
static float rd_f32(rd_t *r) {
    float v;
    rd_raw(r, &v, 4);
    return v;
}
static double rd_f64(rd_t *r) {
    double v;
    rd_raw(r, &v, 8);
    return v;
}

/* ------------------------------------------------------------------ */
/* Rust `as` cast semantics (saturating, NaN -> 0)                    */
/* ------------------------------------------------------------------ */
static int64_t as_i64(double x) {
    if (x != x) return 0;
    if (x >= 9223372036854775808.0) return INT64_MAX;
    if (x <= -9223372036854775808.0) return INT64_MIN;
    return (int64_t)x;
}
static int32_t as_i32(double x) {
    if (x != x) return 0;
    if (x >= 2147483647.0) return INT32_MAX;
    if (x <= -2147483648.0) return INT32_MIN;
    return (int32_t)x;
}
static int16_t as_i16(double x) {
    if (x != x) return 0;
    if (x >= 32767.0) return INT16_MAX;
    if (x <= -32768.0) return INT16_MIN;
    return (int16_t)x;
}
static uint8_t as_u8(double x) {
    if (x != x) return 0;
    if (x >= 255.0) return 255;
    if (x <= 0.0) return 0;
    return (uint8_t)x;
}

/* ------------------------------------------------------------------ */
/* utils/mod.rs                                                       */
/* ------------------------------------------------------------------ */
/* utils/mod.rs:24-29 */
API uint64_t atsc_oracle_prev_power_of_two(uint64_t n) {
    int hb = 63 - __builtin_clzll(n | 1);
    return (1ull << hb) & n;
}
/* utils/mod.rs:41-49 */
static int is_decomposable(uint64_t n) {
    while (n % 2 == 0) n /= 2;
    while (n % 3 == 0) n /= 3;
    return n == 1;
}
/* utils/mod.rs:32-38 */
API uint64_t atsc_oracle_next_size(uint64_t n) {
    n += 1;
    while (!is_decomposable(n)) n += 1;
    return n;
}
/* utils/mod.rs:61-64; 10i32.pow(d) as f64, multiply, round (half away), divide */
static double pow10i(uint32_t d) {
    int32_t y = 1;
    for (uint32_t i = 0; i < d; i++) y *= 10;
    return (double)y;
}
API double atsc_oracle_round_f64(double x, uint32_t decimals) {
    double y = pow10i(decimals);
    return round(x * y) / y;
}
/* utils/mod.rs:66-74 */
API double atsc_oracle_round_and_limit_f64(double x, double mn, double mx, uint32_t decimals) {
    double y = pow10i(decimals);
    double out = round(x * y) / y;
    if (out < mn) return mn;
    if (out > mx) return mx;
    return out;
}

/* utils/error.rs:104-116 -- MAPE, sequential left-to-right f64 sum */
API double atsc_oracle_mape(const double *original, const double *generated, uint64_t n) {
    double s = 0.0;
    for (uint64_t i = 0; i < n; i++) s += fabs((generated[i] - original[i]) / original[i]);
    return s / (double)n;
}

/* ------------------------------------------------------------------ */
/* optimizer/utils.rs                                                 */
/* ------------------------------------------------------------------ */
enum { BD_F64 = 0, BD_I32 = 1, BD_I16 = 2, BD_U8 = 3 };

typedef struct {
    double max, min, mean;
    uint64_t max_loc, min_loc;
    int bitdepth;
    int fractional;
} stats_t;

/* optimizer/utils.rs:115-160 -- exact integer / fraction split.
 * Returns the i64 integer part; *frac_nz = fraction != 0.0 */
static int64_t split_n(double x, int *frac_nz) {
    uint64_t bits;
    memcpy(&bits, &x, 8);
    int is_negative = ((int64_t)bits) < 0;
    int exponent = (int)((bits >> 52) & 0x7FF);
    uint64_t m = (bits & ((1ull << 52) - 1)) | (1ull << 52);
    int64_t mantissa = is_negative ? -(int64_t)m : (int64_t)m;
    int shl = exponent + (64 - 53 - 1023 + 1);
    if (shl <= 0) {
        int shr = -shl;
        if (shr < 64) {
            uint64_t f = ((uint64_t)mantissa) >> shr;
            *frac_nz = f != 0;
            return 0;
        }
        *frac_nz = 0;
        return 0;
    } else if (shl < 64) {
        int64_t i = mantissa >> (64 - shl);
        uint64_t f = ((uint64_t)mantissa) << shl;
        *frac_nz = f != 0;
        return i;
    } else if (shl < 128) {
        int64_t i = (int64_t)(((uint64_t)mantissa) << (shl - 64));
        *frac_nz = 0;
        return i;
    }
    *frac_nz = 0;
    return 0;
}

/* optimizer/utils.rs:91-113 */
static int bitdepth_of(int64_t max_int, int64_t min_int) {
    int bd = max_int <= 255 ? 8 : max_int <= 32767 ? 16 : max_int <= 2147483647ll ? 32 : 64;
    int bs = (min_int >= 0 && min_int <= 255) ? 8
             : min_int >= -32768             ? 16
             : min_int >= -2147483648ll      ? 32
                                             : 64;
    int b = bd > bs ? bd : bs;
    return b == 8 ? BD_U8 : b == 16 ? BD_I16 : b == 32 ? BD_I32 : BD_F64;
}

/* optimizer/utils.rs:39-89 */
static stats_t data_stats(const double *data, uint64_t n) {
    stats_t s;
    s.min = s.max = data[0];
    s.min_loc = s.max_loc = 0;
    s.fractional = 0;
    s.mean = 0.0;
    s.bitdepth = BD_F64;
    for (uint64_t i = 0; i < n; i++) {
        double v = data[i];
        s.mean += v;
        int fnz;
        (void)split_n(v, &fnz);
        if (fnz) s.fractional = 1;
        if (v > s.max) {
            s.max = v;
            s.max_loc = i;
        }
        if (v < s.min) {
            s.min = v;
            s.min_loc = i;
        }
    }
    s.mean /= (double)n;
    int f;
    int64_t max_int = split_n(s.max, &f);
    int64_t min_int = split_n(s.min, &f);
    if (!s.fractional) s.bitdepth = bitdepth_of(max_int, min_int);
    return s;
}

API void atsc_oracle_stats(const double *data, uint64_t n, double *out_min, double *out_max,
                           double *out_mean, uint64_t *min_loc, uint64_t *max_loc, int *bitdepth,
                           int *fractional) {
    stats_t s = data_stats(data, n);
    *out_min = s.min;
    *out_max = s.max;
    *out_mean = s.mean;
    *min_loc = s.min_loc;
    *max_loc = s.max_loc;
    *bitdepth = s.bitdepth;
    *fractional = s.fractional;
}

/* ------------------------------------------------------------------ */
/* compressor ids (compressor/mod.rs:34-44 == bincode variant index)  */
/* ------------------------------------------------------------------ */
enum { C_NOOP = 0, C_FFT = 1, C_IDW = 2, C_CONSTANT = 3, C_POLY = 4, C_AUTO = 5, C_RLE = 6 };

typedef struct {
    buf_t bytes;
    double error;
} result_t;

/* ------------------------------------------------------------------ */
/* constant.rs                                                        */
/* ------------------------------------------------------------------ */
/* constant.rs:37-64 (Encode), :135-139 */
static result_t constant_compressor(const double *data, uint64_t n, stats_t st) {
    (void)data;
    (void)n;
    result_t r = {{0}, 0.0};
    buf_u8(&r.bytes, 30);
    enc_varint(&r.bytes, (uint64_t)st.bitdepth);
    switch (st.bitdepth) {
    case BD_U8: buf_u8(&r.bytes, as_u8(st.min)); break;
    case BD_I16: enc_zigzag(&r.bytes, as_i16(st.min)); break;
    case BD_I32: enc_zigzag(&r.bytes, as_i32(st.min)); break;
    default: enc_f64(&r.bytes, st.min); break;
    }
    return r;
}
/* constant.rs:67-101, :129-132 */
static int constant_to_data(uint64_t n, const uint8_t *p, size_t len, double *out) {
    rd_t r = {p, len, 0, 0};
    (void)rd_u8(&r);
    uint64_t bd = rd_varint(&r);
    double c;
    switch (bd) {
    case BD_U8: c = (double)rd_u8(&r); break;
    case BD_I16: c = (double)(int16_t)rd_zigzag(&r); break;
    case BD_I32: c = (double)(int32_t)rd_zigzag(&r); break;
    case BD_F64: c = rd_f64(&r); break;
    default: return -1;
    }
    if (r.err) return -1;
    for (uint64_t i = 0; i < n; i++) out[i] = c;
    return 0;
}

/* ------------------------------------------------------------------ */
/* noop.rs                                                            */
/* ------------------------------------------------------------------ */
/* noop.rs:37-43, :72-77 */
static buf_t noop_compress(const double *data, uint64_t n) {
    buf_t b = {0};
    buf_u8(&b, 250);
    enc_varint(&b, n);
    for (uint64_t i = 0; i < n; i++) enc_zigzag(&b, as_i64(round(data[i])));
    return b;
}
/* noop.rs:79-83 -- returns number of samples decoded (the stored vector length) */
static int64_t noop_to_data(const uint8_t *p, size_t len, double *out, uint64_t out_cap) {
    rd_t r = {p, len, 0, 0};
    (void)rd_u8(&r);
    uint64_t n = rd_varint(&r);
    if (r.err || n > out_cap) return -1;
    for (uint64_t i = 0; i < n; i++) out[i] = (double)rd_zigzag(&r);
    return r.err ? -1 : (int64_t)n;
}

/* ------------------------------------------------------------------ */
/* rle.rs                                                             */
/* ------------------------------------------------------------------ */
typedef struct {
    uint64_t key; /* f64::to_bits */
    uint64_t idx; /* run start */
} run_t;

static int run_cmp(const void *a, const void *b) {
    const run_t *x = (const run_t *)a, *y = (const run_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    if (x->idx != y->idx) return x->idx < y->idx ? -1 : 1;
    return 0;
}

/* rle.rs:142-189 (IndexRLE::new) + :40-67 (Encode).  The BTreeMap<u64,Vec<usize>>
 * is restated as "collect runs, sort by (bits, start)": same iteration order. */
static result_t rle_compressor(const double *data, uint64_t n, stats_t st) {
    result_t r = {{0}, 0.0};
    run_t *runs = (run_t *)malloc(sizeof(run_t) * (n ? n : 1));
    uint64_t nr = 0, cur = 0;
    for (uint64_t i = 0; i < n; i++) {
        double v = data[i];
        if (i + 1 >= n || data[i + 1] != v) {
            memcpy(&runs[nr].key, &v, 8);
            runs[nr].idx = cur;
            nr++;
            cur = i + 1;
        }
    }
    qsort(runs, nr, sizeof(run_t), run_cmp);
    uint64_t groups = 0;
    for (uint64_t i = 0; i < nr; i++)
        if (i == 0 || runs[i].key != runs[i - 1].key) groups++;
    buf_u8(&r.bytes, 60);
    enc_varint(&r.bytes, (uint64_t)st.bitdepth);
    enc_varint(&r.bytes, groups);
    uint64_t i = 0;
    while (i < nr) {
        uint64_t j = i;
        while (j < nr && runs[j].key == runs[i].key) j++;
        double v;
        memcpy(&v, &runs[i].key, 8);
        switch (st.bitdepth) {
        case BD_U8: buf_u8(&r.bytes, as_u8(v)); break;
        case BD_I16: enc_zigzag(&r.bytes, as_i16(v)); break;
        case BD_I32: enc_zigzag(&r.bytes, as_i32(v)); break;
        default: enc_f64(&r.bytes, v); break;
        }
        enc_varint(&r.bytes, j - i);
        for (uint64_t k = i; k < j; k++) enc_varint(&r.bytes, runs[k].idx);
        i = j;
    }
    free(runs);
    return r;
}

typedef struct {
    uint64_t idx;
    double v;
} iv_t;
static int iv_cmp(const void *a, const void *b) {
    const iv_t *x = (const iv_t *)a, *y = (const iv_t *)b;
    return x->idx < y->idx ? -1 : x->idx > y->idx ? 1 : 0;
}
/* rle.rs:70-110 (Decode), :204-236 (to_data) */
static int rle_to_data(uint64_t n, const uint8_t *p, size_t len, double *out) {
    rd_t r = {p, len, 0, 0};
    (void)rd_u8(&r);
    uint64_t bd = rd_varint(&r);
    uint64_t groups = rd_varint(&r);
    if (r.err || bd > 3) return -1;
    size_t cap = 64, cnt = 0;
    iv_t *fl = (iv_t *)malloc(sizeof(iv_t) * cap);
    for (uint64_t g = 0; g < groups && !r.err; g++) {
        double v;
        switch (bd) {
        case BD_U8: v = (double)rd_u8(&r); break;
        case BD_I16: v = (double)(int16_t)rd_zigzag(&r); break;
        case BD_I32: v = (double)(int32_t)rd_zigzag(&r); break;
        default: v = rd_f64(&r); break;
        }
        uint64_t c = rd_varint(&r);
        for (uint64_t k = 0; k < c && !r.err; k++) {
            if (cnt == cap) {
                cap *= 2;
                fl = (iv_t *)realloc(fl, sizeof(iv_t) * cap);
            }
            fl[cnt].idx = rd_varint(&r);
            fl[cnt].v = v;
            cnt++;
        }
    }
    if (r.err) {
        free(fl);
        return -1;
    }
    qsort(fl, cnt, sizeof(iv_t), iv_cmp);
    for (uint64_t i = 0; i < n; i++) out[i] = 0.0;
    for (size_t i = 0; i < cnt; i++) {
        uint64_t s = fl[i].idx;
        uint64_t e = (i + 1 < cnt) ? fl[i + 1].idx : n;
        /* data.iter_mut().take(end).skip(start) */
        if (e > n) e = n;
        for (uint64_t k = s; k < e; k++) out[k] = fl[i].v;
    }
    free(fl);
    return 0;
}

/* ------------------------------------------------------------------ */
/* polynomial.rs (Polynomial + IDW)                                   */
/* ------------------------------------------------------------------ */
typedef struct {
    int id; /* 0 polynomial, 1 idw */
    double *pts;
    uint64_t npts;
    double min, max;
    uint8_t step;
    int has_error;
    double error;
    int bitdepth;
} poly_t;

/* polynomial.rs:329-340 */
static uint64_t poly_positions(uint64_t frame_size, uint64_t step, uint64_t **out_pos) {
    uint64_t cap = frame_size / (step ? step : 1) + 2;
    uint64_t *pos = (uint64_t *)malloc(sizeof(uint64_t) * cap);
    uint64_t k = 0;
    for (uint64_t v = 0; v < frame_size; v += step) pos[k++] = v;
    if (k == 0 || pos[k - 1] != frame_size - 1) pos[k++] = frame_size - 1;
    *out_pos = pos;
    return k;
}

/* splines 4.3.1 cubic_hermite on (t, value) pairs x(before a), a, b, y(after b).
 * Pinned bit-exactly by the demo HTML polyData arrays (tests/test_oracle_golden.py). */
static double cubic_hermite(double t, double xt, double xv, double at, double av, double bt,
                            double bv, double yt, double yv) {
    double two_t = t * 2.0;
    double three_t = t * 3.0;
    double t2 = t * t;
    double t3 = t2 * t;
    double two_t3 = t2 * two_t;
    double two_t2 = t * two_t;
    double three_t2 = t * three_t;
    double m0 = (bv - xv) / (bt - xt) * (bt - at);
    double m1 = (yv - av) / (yt - at) * (bt - at);
    return av * (two_t3 - three_t2 + 1.0) + m0 * (t3 - two_t2 + t) + bv * (three_t2 - two_t3) +
           m1 * (t3 - t2);
}

/* polynomial.rs:342-373 */
static void polynomial_to_data(const poly_t *p, uint64_t frame_size, double *out) {
    uint64_t *pos;
    uint64_t np = poly_positions(frame_size, p->step, &pos);
    uint64_t K = np < p->npts ? np : p->npts; /* zip stops at the shorter */
    uint64_t seg = 0;
    double prev = p->min;
    for (uint64_t x = 0; x < frame_size; x++) {
        double t = (double)x;
        double v;
        int have = 0;
        if (K >= 2 && t >= (double)pos[0] && t < (double)pos[K - 1]) {
            /* search_lower_cp: key i with pos[i] <= t < pos[i+1] */
            while (seg + 1 < K - 1 && (double)pos[seg + 1] <= t) seg++;
            uint64_t i = seg;
            int catmull = (i > 0 && K - i > 2);
            double at = (double)pos[i], bt = (double)pos[i + 1];
            double nt = (t - at) / (bt - at);
            if (!catmull) {
                v = p->pts[i] * (1.0 - nt) + p->pts[i + 1] * nt;
                have = 1;
            } else if (!(i == 0 || i >= K - 2)) {
                v = cubic_hermite(nt, (double)pos[i - 1], p->pts[i - 1], at, p->pts[i], bt,
                                  p->pts[i + 1], (double)pos[i + 2], p->pts[i + 2]);
                have = 1;
            }
        }
        if (!have && K >= 1) {
            /* clamped_sample fallbacks */
            if (t <= (double)pos[0]) {
                v = p->pts[0];
                have = 1;
            } else if (t >= (double)pos[K - 1]) {
                v = p->pts[K - 1];
                have = 1;
            }
        }
        if (!have) v = prev; /* unwrap_or(prev) */
        prev = v;
        out[x] = atsc_oracle_round_and_limit_f64(v, p->min, p->max, 5);
    }
    free(pos);
}

/* inverse_distance_weight 0.1.1, power 2, all points; pinned by
 * polynomial.rs:545-567 and the demo HTML idwData arrays. */
static int g_idw_variant = 0; /* test hook: alternative summation orders */
API void atsc_oracle_set_idw_variant(int v) { g_idw_variant = v; }

static double idw_eval(const uint64_t *pos, const double *vals, uint64_t K, double x, double *w) {
    for (uint64_t j = 0; j < K; j++) {
        double d = fabs(x - (double)pos[j]);
        if (d == 0.0) return vals[j];
        w[j] = 1.0 / pow(d, 2.0);
    }
    double S = 0.0;
    for (uint64_t j = 0; j < K; j++) S += w[j];
    double acc = 0.0;
    switch (g_idw_variant) {
    default:
    case 0:
        for (uint64_t j = 0; j < K; j++) acc += (w[j] / S) * vals[j];
        return acc;
    case 1:
        for (uint64_t j = 0; j < K; j++) acc += w[j] * vals[j];
        return acc / S;
    case 2:
        for (uint64_t j = 0; j < K; j++) acc += w[j] * vals[j] / S;
        return acc;
    case 3:
        for (uint64_t j = 0; j < K; j++) acc += vals[j] * w[j] / S;
        return acc;
    }
}

/* polynomial.rs:375-393 */
static void idw_to_data(const poly_t *p, uint64_t frame_size, double *out) {
    uint64_t *pos;
    uint64_t np = poly_positions(frame_size, p->step, &pos);
    uint64_t K = np < p->npts ? np : p->npts;
    double *w = (double *)malloc(sizeof(double) * (K ? K : 1));
    for (uint64_t x = 0; x < frame_size; x++) {
        double v = idw_eval(pos, p->pts, K, (double)x, w);
        out[x] = atsc_oracle_round_and_limit_f64(v, p->min, p->max, 5);
    }
    free(w);
    free(pos);
}

/* polynomial.rs:395-404 */
static void poly_to_data(const poly_t *p, uint64_t frame_size, double *out) {
    if (p->max == p->min) {
        for (uint64_t i = 0; i < frame_size; i++) out[i] = p->max;
        return;
    }
    if (p->id == 1)
        idw_to_data(p, frame_size, out);
    else
        polynomial_to_data(p, frame_size, out);
}

/* polynomial.rs:279-305 */
static void poly_compress_hinted(poly_t *p, const double *data, uint64_t n, uint64_t points) {
    if (p->max == p->min) return;
    uint64_t step = n / points;
    if (step < 1) step = 1;
    uint64_t *pos;
    uint64_t k = poly_positions(n, step, &pos); /* (0..n).step_by(step) + last */
    free(p->pts);
    p->pts = (double *)malloc(sizeof(double) * k);
    for (uint64_t i = 0; i < k; i++) p->pts[i] = data[pos[i]];
    p->npts = k;
    p->step = (uint8_t)step; /* `step as u8` */
    free(pos);
}

/* polynomial.rs:209-277 */
static void poly_compress_bounded(poly_t *p, const double *data, uint64_t n, double max_err,
                                  int *iters_out) {
    *iters_out = 0;
    if (p->max == p->min) return;
    uint64_t baseline = (3 >= n / 100) ? 3 : n / 100;
    double cur = max_err + 1.0;
    uint64_t jump = 0;
    int it = 0;
    double target = atsc_oracle_round_f64(max_err, 3);
    double *out = (double *)malloc(sizeof(double) * n);
    while (target < atsc_oracle_round_f64(cur, 4)) {
        it++;
        poly_compress_hinted(p, data, n, baseline + jump);
        if (p->id == 1)
            idw_to_data(p, n, out);
        else
            polynomial_to_data(p, n, out);
        cur = atsc_oracle_mape(data, out, n);
        if (it >= 1 && it <= 17) {
            uint64_t j = n / 10;
            jump += j > 1 ? j : 1;
        } else if (it >= 18 && it <= 22) {
            uint64_t j = n / 100;
            jump += j > 1 ? j : 1;
        } else if (target > atsc_oracle_round_f64(cur, 4)) {
            break;
        } else {
            poly_compress_hinted(p, data, n, n);
            cur = 0.0;
            break;
        }
        if (p->npts == n) {
            cur = 0.0;
            break;
        }
    }
    free(out);
    p->has_error = 1;
    p->error = cur;
    *iters_out = it;
}

/* polynomial.rs:54-87 */
static void poly_encode(const poly_t *p, buf_t *b) {
    enc_varint(b, (uint64_t)p->id);
    enc_varint(b, (uint64_t)p->bitdepth);
    enc_varint(b, p->npts);
    for (uint64_t i = 0; i < p->npts; i++) {
        switch (p->bitdepth) {
        case BD_U8: buf_u8(b, as_u8(p->pts[i])); break;
        case BD_I16: enc_zigzag(b, as_i16(p->pts[i])); break;
        case BD_I32: enc_zigzag(b, as_i32(p->pts[i])); break;
        default: enc_f64(b, p->pts[i]); break;
        }
    }
    enc_f64(b, p->min);
    enc_f64(b, p->max);
    buf_u8(b, p->step);
}

/* polynomial.rs:89-131 */
static int poly_decode(poly_t *p, const uint8_t *bytes, size_t len) {
    rd_t r = {bytes, len, 0, 0};
    memset(p, 0, sizeof(*p));
    p->id = (int)rd_varint(&r);
    p->bitdepth = (int)rd_varint(&r);
    uint64_t k = rd_varint(&r);
    if (r.err || p->bitdepth > 3 || p->id > 1 || k > len) return -1;
    p->pts = (double *)malloc(sizeof(double) * (k ? k : 1));
    p->npts = k;
    for (uint64_t i = 0; i < k; i++) {
        switch (p->bitdepth) {
        case BD_U8: p->pts[i] = (double)rd_u8(&r); break;
        case BD_I16: p->pts[i] = (double)(int16_t)rd_zigzag(&r); break;
        case BD_I32: p->pts[i] = (double)(int32_t)rd_zigzag(&r); break;
        default: p->pts[i] = rd_f64(&r); break;
        }
    }
    p->min = rd_f64(&r);
    p->max = rd_f64(&r);
    p->step = rd_u8(&r);
    if (r.err) {
        free(p->pts);
        p->pts = NULL;
        return -1;
    }
    return 0;
}

static poly_t poly_new(stats_t st, int ptype) {
    poly_t p;
    memset(&p, 0, sizeof(p));
    p.id = ptype;
    p.min = st.min;
    p.max = st.max;
    p.step = 1;
    p.bitdepth = st.bitdepth;
    return p;
}

/* polynomial.rs:407-413 */
static buf_t polynomial_unbounded(const double *data, uint64_t n, int ptype) {
    stats_t st = data_stats(data, n);
    poly_t p = poly_new(st, ptype);
    uint64_t points = (3 >= n / 100) ? 3 : n / 100;
    poly_compress_hinted(&p, data, n, points);
    buf_t b = {0};
    poly_encode(&p, &b);
    free(p.pts);
    return b;
}

/* polynomial.rs:415-425 */
static __thread int g_last_poly_iters = 0;
static result_t polynomial_allowed_error(const double *data, uint64_t n, double allowed,
                                         int ptype) {
    stats_t st = data_stats(data, n);
    poly_t p = poly_new(st, ptype);
    int it;
    poly_compress_bounded(&p, data, n, allowed, &it);
    g_last_poly_iters = it;
    result_t r = {{0}, p.has_error ? p.error : 0.0};
    poly_encode(&p, &r.bytes);
    free(p.pts);
    return r;
}

/* polynomial.rs:427-430 */
static int poly_payload_to_data(uint64_t n, const uint8_t *bytes, size_t len, double *out) {
    poly_t p;
    if (poly_decode(&p, bytes, len)) return -1;
    if (p.step == 0 && p.max != p.min) { /* step_by(0) panics in the reference */
        free(p.pts);
        return -1;
    }
    poly_to_data(&p, n, out);
    free(p.pts);
    return 0;
}

/* ------------------------------------------------------------------ */
/* f32 complex DFT (restates rustfft's contract: unnormalised,        */
/* forward = exp(-2 pi i jk/n), any length)                           */
/* ------------------------------------------------------------------ */
typedef struct {
    float re, im;
} cf32;

typedef struct {
    int n;
    int inverse;
    cf32 *tw; /* tw[j] = exp(-/+ 2 pi i j / n) */
    int nf;
    int fac[64]; /* radix sequence */
    cf32 *scratch;
} fft_plan;

static fft_plan *fft_plan_new(int n, int inverse) {
    fft_plan *p = (fft_plan *)calloc(1, sizeof(fft_plan));
    p->n = n;
    p->inverse = inverse;
    p->tw = (cf32 *)malloc(sizeof(cf32) * (n > 0 ? n : 1));
    for (int j = 0; j < n; j++) {
        double a = 2.0 * M_PI * (double)j / (double)n;
        p->tw[j].re = (float)cos(a);
        p->tw[j].im = (float)(inverse ? sin(a) : -sin(a));
    }
    int m = n;
    while (m % 4 == 0) {
        p->fac[p->nf++] = 4;
        m /= 4;
    }
    while (m % 2 == 0) {
        p->fac[p->nf++] = 2;
        m /= 2;
    }
    for (int f = 3; m > 1; f += 2) {
        while (m % f == 0) {
            p->fac[p->nf++] = f;
            m /= f;
        }
        if (f * f > m && m > 1) {
            p->fac[p->nf++] = m;
            m = 1;
        }
    }
    p->scratch = (cf32 *)malloc(sizeof(cf32) * 128);
    return p;
}
static void fft_plan_free(fft_plan *p) {
    if (!p) return;
    free(p->tw);
    free(p->scratch);
    free(p);
}

static inline cf32 cmul(cf32 a, cf32 b) {
    cf32 r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
    return r;
}
static inline cf32 cadd(cf32 a, cf32 b) {
    cf32 r = {a.re + b.re, a.im + b.im};
    return r;
}
static inline cf32 csub(cf32 a, cf32 b) {
    cf32 r = {a.re - b.re, a.im - b.im};
    return r;
}

/* recursive decimation-in-time mixed radix (textbook Cooley-Tukey) */
static void fft_rec(const fft_plan *p, cf32 *out, const cf32 *in, int n, int stride, int fi) {
    if (n == 1) {
        out[0] = in[0];
        return;
    }
    int r = p->fac[fi];
    int m = n / r;
    for (int s = 0; s < r; s++) fft_rec(p, out + s * m, in + (size_t)s * stride, m, stride * r, fi + 1);
    int N = p->n;
    int tws = N / n; /* twiddle stride: w_n^j = tw[j * tws] */
    if (r == 2) {
        for (int k = 0; k < m; k++) {
            cf32 a = out[k];
            cf32 b = cmul(out[k + m], p->tw[(size_t)k * tws]);
            out[k] = cadd(a, b);
            out[k + m] = csub(a, b);
        }
    } else if (r == 4) {
        for (int k = 0; k < m; k++) {
            cf32 a = out[k];
            cf32 b = cmul(out[k + m], p->tw[(size_t)k * tws]);
            cf32 c = cmul(out[k + 2 * m], p->tw[(size_t)2 * k * tws]);
            cf32 d = cmul(out[k + 3 * m], p->tw[(size_t)3 * k * tws]);
            cf32 t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = csub(b, d);
            /* forward: -i * t3 ; inverse: +i * t3 */
            cf32 jt3;
            if (p->inverse) {
                jt3.re = -t3.im;
                jt3.im = t3.re;
            } else {
                jt3.re = t3.im;
                jt3.im = -t3.re;
            }
            out[k] = cadd(t0, t2);
            out[k + m] = cadd(t1, jt3);
            out[k + 2 * m] = csub(t0, t2);
            out[k + 3 * m] = csub(t1, jt3);
        }
    } else if (r == 3) {
        cf32 w3 = p->tw[(size_t)m * tws]; /* exp(-/+ 2 pi i / 3) */
        for (int k = 0; k < m; k++) {
            cf32 a = out[k];
            cf32 b = cmul(out[k + m], p->tw[(size_t)k * tws]);
            cf32 c = cmul(out[k + 2 * m], p->tw[(size_t)2 * k * tws]);
            cf32 s = cadd(b, c), d = csub(b, c);
            cf32 t = {a.re + w3.re * s.re, a.im + w3.re * s.im};
            cf32 u = {-w3.im * d.im, w3.im * d.re};
            out[k] = cadd(a, s);
            out[k + m] = cadd(t, u);
            out[k + 2 * m] = csub(t, u);
        }
    } else {
        cf32 *sc = p->scratch;
        cf32 *tmp = (r <= 128) ? sc : (cf32 *)malloc(sizeof(cf32) * r);
        for (int k = 0; k < m; k++) {
            for (int s = 0; s < r; s++) tmp[s] = out[k + s * m];
            for (int q = 0; q < r; q++) {
                int kk = k + q * m;
                cf32 acc = tmp[0];
                size_t tw = 0;
                for (int s = 1; s < r; s++) {
                    tw += (size_t)kk * tws;
                    tw %= (size_t)N;
                    acc = cadd(acc, cmul(tmp[s], p->tw[tw]));
                }
                out[kk] = acc;
            }
        }
        if (tmp != sc) free(tmp);
    }
}

static void fft_process(const fft_plan *p, cf32 *buf) {
    if (p->n <= 1) return;
    cf32 *tmp = (cf32 *)malloc(sizeof(cf32) * p->n);
    memcpy(tmp, buf, sizeof(cf32) * p->n);
    fft_rec(p, buf, tmp, p->n, 1, 0);
    free(tmp);
}

/* ------------------------------------------------------------------ */
/* fft.rs                                                             */
/* ------------------------------------------------------------------ */
typedef struct {
    uint16_t pos;
    float re, im;
} fpoint_t;

typedef struct {
    fpoint_t *f;
    uint64_t nf;
    float max_value, min_value;
    int has_error;
    double error;
} fftc_t;

/* fft.rs:66-106: ordering on Complex<f32>::norm() == hypotf */
static inline float fp_norm(const fpoint_t *a) { return hypotf(a->re, a->im); }
static inline int fp_le(const fpoint_t *a, const fpoint_t *b) { return fp_norm(a) <= fp_norm(b); }
static inline int fp_ge(const fpoint_t *a, const fpoint_t *b) { return fp_norm(a) >= fp_norm(b); }
static inline int fp_lt(const fpoint_t *a, const fpoint_t *b) { return fp_norm(a) < fp_norm(b); }

/* std::collections::BinaryHeap (max-heap) restated so equal-norm tie order
 * follows the reference's pop order. Norms are cached for speed. */
typedef struct {
    fpoint_t p;
    float n;
} hnode_t;

static void heap_sift_down_range(hnode_t *d, size_t pos, size_t end) {
    hnode_t elt = d[pos];
    size_t child = 2 * pos + 1;
    size_t lim = end >= 2 ? end - 2 : 0;
    while (child <= lim && end >= 2) {
        child += (d[child].n <= d[child + 1].n) ? 1 : 0;
        if (elt.n >= d[child].n) {
            d[pos] = elt;
            return;
        }
        d[pos] = d[child];
        pos = child;
        child = 2 * pos + 1;
    }
    if (child == end - 1 && elt.n < d[child].n) {
        d[pos] = d[child];
        pos = child;
    }
    d[pos] = elt;
}
static void heap_rebuild(hnode_t *d, size_t len) {
    size_t n = len / 2;
    while (n > 0) {
        n--;
        heap_sift_down_range(d, n, len);
    }
}
static void heap_sift_down_to_bottom(hnode_t *d, size_t len) {
    size_t end = len, start = 0, pos = 0;
    hnode_t elt = d[pos];
    size_t child = 1;
    size_t lim = end >= 2 ? end - 2 : 0;
    while (child <= lim && end >= 2) {
        child += (d[child].n <= d[child + 1].n) ? 1 : 0;
        d[pos] = d[child];
        pos = child;
        child = 2 * pos + 1;
    }
    if (child == end - 1) {
        d[pos] = d[child];
        pos = child;
    }
    /* sift_up(start, pos) */
    while (pos > start) {
        size_t parent = (pos - 1) / 2;
        if (elt.n <= d[parent].n) break;
        d[pos] = d[parent];
        pos = parent;
    }
    d[pos] = elt;
}
static int heap_pop(hnode_t *d, size_t *len, hnode_t *out) {
    if (*len == 0) return 0;
    hnode_t item = d[*len - 1];
    (*len)--;
    if (*len > 0) {
        hnode_t t = d[0];
        d[0] = item;
        item = t;
        heap_sift_down_to_bottom(d, *len);
    }
    *out = item;
    return 1;
}

/* fft.rs:231-257 */
static void fft_trim(fftc_t *c, const cf32 *buffer, uint64_t blen, uint64_t max_freq) {
    free(c->f);
    c->f = (fpoint_t *)malloc(sizeof(fpoint_t) * (max_freq ? max_freq : 1));
    c->nf = 0;
    if (max_freq == 1) {
        c->f[0].pos = 0;
        c->f[0].re = buffer[0].re;
        c->f[0].im = buffer[0].im;
        c->nf = 1;
        return;
    }
    hnode_t *h = (hnode_t *)malloc(sizeof(hnode_t) * (blen ? blen : 1));
    for (uint64_t i = 0; i < blen; i++) {
        h[i].p.pos = (uint16_t)i; /* `pos as u16` wraps */
        h[i].p.re = buffer[i].re;
        h[i].p.im = buffer[i].im;
        h[i].n = fp_norm(&h[i].p);
    }
    size_t len = blen;
    heap_rebuild(h, len);
    for (uint64_t k = 0; k < max_freq; k++) {
        hnode_t it;
        if (heap_pop(h, &len, &it)) {
            if (it.p.im == 0.0f && it.p.re == 0.0f) break;
            c->f[c->nf++] = it.p;
        }
    }
    free(h);
}

/* fft.rs:401-422 */
static void get_mirrored_freqs(const fftc_t *c, cf32 *data, uint64_t len) {
    memset(data, 0, sizeof(cf32) * len);
    for (uint64_t i = 0; i < c->nf; i++) {
        uint64_t pos = c->f[i].pos;
        if (pos >= len) continue; /* reference would panic (index out of bounds) */
        data[pos].re = c->f[i].re;
        data[pos].im = c->f[i].im;
        if (pos == 0) continue;
        data[len - pos].re = c->f[i].re;
        data[len - pos].im = c->f[i].im * -1.0f;
    }
}

/* fft.rs:208-218 */
static inline double fft_round(const fftc_t *c, float x) {
    double y = 100000.0;
    double out = round((double)x * y) / y;
    if (out > (double)c->max_value) return (double)c->max_value;
    if (out < (double)c->min_value) return (double)c->min_value;
    return out;
}

/* fft.rs:184-204 */
static double *gibbs_sizing(const double *data, uint64_t n, uint64_t *out_len) {
    uint64_t ns = atsc_oracle_next_size(n);
    uint64_t added = ns - n;
    uint64_t prefix = added / 2, suffix = added - prefix;
    double *r = (double *)malloc(sizeof(double) * ns);
    uint64_t k = 0;
    for (uint64_t i = 0; i < prefix; i++) r[k++] = data[0];
    for (uint64_t i = 0; i < n; i++) r[k++] = data[i];
    for (uint64_t i = 0; i < suffix; i++) r[k++] = data[n - 1];
    *out_len = ns;
    return r;
}

/* fft.rs:162-171 */
static fftc_t fftc_new(double mn, double mx) {
    fftc_t c;
    memset(&c, 0, sizeof(c));
    c.max_value = (float)mx;
    c.min_value = (float)mn;
    return c;
}

static __thread int g_last_fft_iters = 0;

/* fft.rs:288-362 */
static void fft_compress_bounded(fftc_t *c, const double *data, uint64_t n, double max_err) {
    g_last_fft_iters = 0;
    if (c->max_value == c->min_value) return;
    uint64_t max_freq = (3 >= n / 100) ? 3 : n / 100;
    double *g_alloc = NULL;
    const double *g = data;
    uint64_t len = n;
    if (n >= 128) {
        g_alloc = gibbs_sizing(data, n, &len);
        g = g_alloc;
    }
    float len_f32 = (float)len;
    cf32 *buffer = (cf32 *)malloc(sizeof(cf32) * len);
    for (uint64_t i = 0; i < len; i++) {
        buffer[i].re = (float)g[i];
        buffer[i].im = 0.0f;
    }
    fft_plan *fwd = fft_plan_new((int)len, 0);
    fft_plan *inv = fft_plan_new((int)len, 1);
    fft_process(fwd, buffer);
    uint64_t size = len / 2 + 1;
    double cur = max_err + 1.0;
    uint64_t jump = 0;
    int it = 0;
    cf32 *idata = (cf32 *)malloc(sizeof(cf32) * len);
    double *out = (double *)malloc(sizeof(double) * len);
    while (as_i32(max_err * 1000.0) < as_i32(cur * 1000.0)) {
        it++;
        fft_trim(c, buffer, size, max_freq + jump);
        get_mirrored_freqs(c, idata, len);
        fft_process(inv, idata);
        for (uint64_t i = 0; i < len; i++) out[i] = fft_round(c, idata[i].re / len_f32);
        cur = atsc_oracle_mape(g, out, len);
        if (it >= 1 && it <= 17) {
            uint64_t j = max_freq / 2;
            jump += j > 1 ? j : 1;
        } else if (it >= 18 && it <= 22) {
            uint64_t j = max_freq / 10;
            jump += j > 1 ? j : 1;
        } else {
            break;
        }
    }
    c->has_error = 1;
    c->error = cur;
    g_last_fft_iters = it;
    free(out);
    free(idata);
    free(buffer);
    free(g_alloc);
    fft_plan_free(fwd);
    fft_plan_free(inv);
}

/* fft.rs:262-282 (compress_hinted) and :366-388 (compress): NO gibbs padding */
static void fft_compress_hinted(fftc_t *c, const double *data, uint64_t n, uint64_t max_freq) {
    if (c->max_value == c->min_value) return;
    cf32 *buffer = (cf32 *)malloc(sizeof(cf32) * n);
    for (uint64_t i = 0; i < n; i++) {
        buffer[i].re = (float)data[i];
        buffer[i].im = 0.0f;
    }
    fft_plan *fwd = fft_plan_new((int)n, 0);
    fft_process(fwd, buffer);
    fft_trim(c, buffer, n / 2 + 1, max_freq);
    fft_plan_free(fwd);
    free(buffer);
}

/* fft.rs:119-130 */
static void fft_encode(const fftc_t *c, buf_t *b) {
    buf_u8(b, 15);
    enc_varint(b, c->nf);
    for (uint64_t i = 0; i < c->nf; i++) {
        enc_varint(b, c->f[i].pos);
        enc_f32(b, c->f[i].re);
        enc_f32(b, c->f[i].im);
    }
    enc_f32(b, c->max_value);
    enc_f32(b, c->min_value);
}

/* fft.rs:132-144 */
static int fft_decode(fftc_t *c, const uint8_t *bytes, size_t len) {
    rd_t r = {bytes, len, 0, 0};
    memset(c, 0, sizeof(*c));
    (void)rd_u8(&r);
    uint64_t k = rd_varint(&r);
    if (r.err || k > len) return -1;
    c->f = (fpoint_t *)malloc(sizeof(fpoint_t) * (k ? k : 1));
    c->nf = k;
    for (uint64_t i = 0; i < k; i++) {
        c->f[i].pos = (uint16_t)rd_varint(&r);
        c->f[i].re = rd_f32(&r);
        c->f[i].im = rd_f32(&r);
    }
    c->max_value = rd_f32(&r);
    c->min_value = rd_f32(&r);
    if (r.err) {
        free(c->f);
        c->f = NULL;
        return -1;
    }
    return 0;
}

/* fft.rs:426-462 */
static int fft_to_data(uint64_t frame_size, const uint8_t *bytes, size_t blen, double *out) {
    fftc_t c;
    if (fft_decode(&c, bytes, blen)) return -1;
    if (c.max_value == c.min_value) {
        for (uint64_t i = 0; i < frame_size; i++) out[i] = (double)c.max_value;
        free(c.f);
        return 0;
    }
    uint64_t prefix = 0, suffix = 0;
    if (frame_size >= 128) {
        uint64_t added = atsc_oracle_next_size(frame_size) - frame_size;
        prefix = added / 2;
        suffix = added - prefix;
    }
    uint64_t glen = frame_size + prefix + suffix;
    cf32 *data = (cf32 *)malloc(sizeof(cf32) * glen);
    get_mirrored_freqs(&c, data, glen);
    fft_plan *inv = fft_plan_new((int)glen, 1);
    fft_process(inv, data);
    float lf = (float)glen;
    for (uint64_t i = 0; i < frame_size; i++) out[i] = fft_round(&c, data[i + prefix].re / lf);
    fft_plan_free(inv);
    free(data);
    free(c.f);
    return 0;
}

/* fft.rs:516-524 */
static result_t fft_compressor(const double *data, uint64_t n, double allowed, stats_t st) {
    fftc_t c = fftc_new(st.min, st.max);
    fft_compress_bounded(&c, data, n, allowed);
    result_t r = {{0}, c.has_error ? c.error : 0.0};
    fft_encode(&c, &r.bytes);
    free(c.f);
    return r;
}

static void minmax_scan(const double *data, uint64_t n, double *mn, double *mx) {
    *mn = *mx = data[0];
    for (uint64_t i = 0; i < n; i++) {
        if (data[i] > *mx) *mx = data[i];
        if (data[i] < *mn) *mn = data[i];
    }
}

/* fft.rs:466-484 (`fft`): max(3, n/100) frequencies, no padding */
static buf_t fft_unbounded(const double *data, uint64_t n) {
    double mn, mx;
    minmax_scan(data, n, &mn, &mx);
    fftc_t c = fftc_new(mn, mx);
    uint64_t max_freq = (3 >= n / 100) ? 3 : n / 100;
    fft_compress_hinted(&c, data, n, max_freq);
    buf_t b = {0};
    fft_encode(&c, &b);
    free(c.f);
    return b;
}

/* ------------------------------------------------------------------ */
/* compressor/mod.rs dispatch                                         */
/* ------------------------------------------------------------------ */
/* compressor/mod.rs:94-107 */
static int get_compress_bounded_results(int comp, const double *data, uint64_t n, double max_error,
                                        result_t *out) {
    stats_t st = data_stats(data, n);
    switch (comp) {
    case C_NOOP:
        out->bytes = noop_compress(data, n);
        out->error = 0.0;
        return 0;
    case C_FFT: *out = fft_compressor(data, n, max_error, st); return 0;
    case C_CONSTANT: *out = constant_compressor(data, n, st); return 0;
    case C_RLE: *out = rle_compressor(data, n, st); return 0;
    case C_POLY: *out = polynomial_allowed_error(data, n, max_error, 0); return 0;
    case C_IDW: *out = polynomial_allowed_error(data, n, max_error, 1); return 0;
    default: return -1; /* todo!() */
    }
}

/* compressor/mod.rs:63-74 */
static int compress_unbounded(int comp, const double *data, uint64_t n, buf_t *out) {
    stats_t st = data_stats(data, n);
    result_t r;
    switch (comp) {
    case C_NOOP: *out = noop_compress(data, n); return 0;
    case C_FFT: *out = fft_unbounded(data, n); return 0;
    case C_CONSTANT:
        r = constant_compressor(data, n, st);
        *out = r.bytes;
        return 0;
    case C_POLY: *out = polynomial_unbounded(data, n, 0); return 0;
    case C_IDW: *out = polynomial_unbounded(data, n, 1); return 0;
    case C_RLE:
        r = rle_compressor(data, n, st);
        *out = r.bytes;
        return 0;
    default: return -1;
    }
}

/* compressor/mod.rs:109-119.  Returns decoded sample count or <0. */
static int64_t decompress_payload(int comp, uint64_t samples, const uint8_t *p, size_t len,
                                  double *out) {
    switch (comp) {
    case C_NOOP: return noop_to_data(p, len, out, samples);
    case C_FFT: return fft_to_data(samples, p, len, out) ? -1 : (int64_t)samples;
    case C_CONSTANT: return constant_to_data(samples, p, len, out) ? -1 : (int64_t)samples;
    case C_POLY:
    case C_IDW: return poly_payload_to_data(samples, p, len, out) ? -1 : (int64_t)samples;
    case C_RLE: return rle_to_data(samples, p, len, out) ? -1 : (int64_t)samples;
    default: return -1;
    }
}

/* ------------------------------------------------------------------ */
/* frame/mod.rs                                                       */
/* ------------------------------------------------------------------ */
static const int64_t COMPRESSION_SPEED[7] = {2147483647, 4096, 2048, 1024, 512, 256, 128};

typedef struct {
    int compressor;
    uint64_t sample_count;
    buf_t data;
    /* diagnostics (not serialised) */
    double cand_err[3];
    uint64_t cand_size[3];
    int cand_valid;
} frame_t;

/* frame/mod.rs:71-149 */
static int compress_best(frame_t *f, const double *data, uint64_t n, float max_error_f32,
                         uint32_t speed) {
    double max_error = (double)max_error_f32;
    f->sample_count = n;
    f->cand_valid = 0;
    uint64_t data_sample = (uint64_t)COMPRESSION_SPEED[speed];
    static const int list[3] = {C_FFT, C_POLY, C_RLE};
    stats_t st = data_stats(data, n);
    result_t r;
    if (st.min == st.max) {
        f->compressor = C_CONSTANT;
        get_compress_bounded_results(C_CONSTANT, data, n, max_error, &r);
        f->data = r.bytes;
        return 0;
    }
    if (n >= data_sample) {
        int best = -1;
        uint64_t best_len = 0;
        for (int i = 0; i < 3; i++) {
            get_compress_bounded_results(list[i], data, data_sample, max_error, &r);
            f->cand_err[i] = r.error;
            f->cand_size[i] = r.bytes.len;
            if (r.error <= max_error && (best < 0 || r.bytes.len < best_len)) {
                best = i;
                best_len = r.bytes.len;
            }
            buf_free(&r.bytes);
        }
        f->cand_valid = 1;
        if (best < 0) return -2; /* unwrap() on None */
        f->compressor = list[best];
        get_compress_bounded_results(f->compressor, data, n, max_error, &r);
        f->data = r.bytes;
        return 0;
    }
    result_t rs[3];
    int all_fail = 1;
    for (int i = 0; i < 3; i++) {
        get_compress_bounded_results(list[i], data, n, max_error, &rs[i]);
        f->cand_err[i] = rs[i].error;
        f->cand_size[i] = rs[i].bytes.len;
        if (rs[i].error <= max_error) all_fail = 0;
    }
    f->cand_valid = 1;
    int best = -1;
    for (int i = 0; i < 3; i++) {
        if (!all_fail && !(rs[i].error <= max_error)) continue;
        if (best < 0 || rs[i].bytes.len < rs[best].bytes.len) best = i;
    }
    f->compressor = list[best];
    f->data = rs[best].bytes;
    for (int i = 0; i < 3; i++)
        if (i != best) buf_free(&rs[i].bytes);
    return 0;
}

/* ------------------------------------------------------------------ */
/* exported per-frame API                                             */
/* ------------------------------------------------------------------ */
static int64_t emit(buf_t *b, uint8_t *out, uint64_t cap) {
    int64_t n = (int64_t)b->len;
    if (b->len > cap) n = -(int64_t)b->len - 100; /* need more room */
    else if (b->len)
        memcpy(out, b->p, b->len);
    buf_free(b);
    return n;
}

/* Compressor::compress (compressor/mod.rs:63) */
API int64_t atsc_oracle_compress(int comp, const double *data, uint64_t n, uint8_t *out,
                                 uint64_t cap) {
    buf_t b = {0};
    if (n == 0 || compress_unbounded(comp, data, n, &b)) return -1;
    return emit(&b, out, cap);
}

/* Compressor::get_compress_bounded_results (compressor/mod.rs:94); also
 * compress_bounded (:76) which returns the same bytes. */
API int64_t atsc_oracle_compress_bounded(int comp, const double *data, uint64_t n,
                                         double max_error, uint8_t *out, uint64_t cap,
                                         double *error, int *iterations) {
    result_t r;
    g_last_fft_iters = g_last_poly_iters = 0;
    if (n == 0 || get_compress_bounded_results(comp, data, n, max_error, &r)) return -1;
    if (error) *error = r.error;
    if (iterations) *iterations = comp == C_FFT ? g_last_fft_iters : g_last_poly_iters;
    return emit(&r.bytes, out, cap);
}

/* CompressorFrame::compress_best (frame/mod.rs:71) */
API int64_t atsc_oracle_compress_best(const double *data, uint64_t n, float max_error,
                                      uint32_t speed, int *compressor, uint8_t *out, uint64_t cap,
                                      double *cand_err3, uint64_t *cand_size3) {
    frame_t f;
    memset(&f, 0, sizeof(f));
    if (n == 0 || speed > 6) return -1;
    int rc = compress_best(&f, data, n, max_error, speed);
    if (rc) return rc;
    *compressor = f.compressor;
    for (int i = 0; i < 3; i++) {
        if (cand_err3) cand_err3[i] = f.cand_valid ? f.cand_err[i] : 0.0;
        if (cand_size3) cand_size3[i] = f.cand_valid ? f.cand_size[i] : 0;
    }
    return emit(&f.data, out, cap);
}

/* Compressor::decompress (compressor/mod.rs:109) */
API int64_t atsc_oracle_decompress(int comp, uint64_t samples, const uint8_t *payload,
                                   uint64_t len, double *out) {
    return decompress_payload(comp, samples, payload, len, out);
}

/* FFT helpers exposed for the golden tests (fft.rs:526-544 `fft_set`) */
API int64_t atsc_oracle_fft_set(const double *data, uint64_t n, uint64_t freqs, uint8_t *out,
                                uint64_t cap) {
    double mn, mx;
    minmax_scan(data, n, &mn, &mx);
    fftc_t c = fftc_new(mn, mx);
    fft_compress_hinted(&c, data, n, freqs);
    buf_t b = {0};
    fft_encode(&c, &b);
    free(c.f);
    return emit(&b, out, cap);
}

API uint64_t atsc_oracle_gibbs_sizing(const double *data, uint64_t n, double *out, uint64_t cap) {
    uint64_t len;
    double *g = gibbs_sizing(data, n, &len);
    if (len <= cap) memcpy(out, g, sizeof(double) * len);
    free(g);
    return len;
}

/* raw f32 DFT, for cross-checking against numpy */
API void atsc_oracle_fft_c32(float *interleaved, int n, int inverse) {
    fft_plan *p = fft_plan_new(n, inverse);
    fft_process(p, (cf32 *)interleaved);
    fft_plan_free(p);
}

/* ------------------------------------------------------------------ */
/* optimizer/mod.rs (planner)                                         */
/* ------------------------------------------------------------------ */
/* optimizer/mod.rs:64-71 */
API uint64_t atsc_oracle_clean_data(const double *in, uint64_t n, double *out) {
    uint64_t k = 0;
    for (uint64_t i = 0; i < n; i++)
        if (!(isnan(in[i]) || isinf(in[i]))) out[k++] = in[i];
    return k;
}
/* optimizer/mod.rs:78-98 */
API uint64_t atsc_oracle_chunk_sizes(uint64_t len, uint64_t *out, uint64_t cap) {
    uint64_t k = 0;
    while (len > 0) {
        uint64_t s;
        if (len >= 131072)
            s = 131072;
        else if (len <= 512)
            s = len;
        else
            s = atsc_oracle_prev_power_of_two(len);
        if (k < cap) out[k] = s;
        k++;
        len -= s;
    }
    return k;
}

/* ------------------------------------------------------------------ */
/* data.rs / header.rs: BRO stream                                    */
/* ------------------------------------------------------------------ */
/* main.rs:130-166 compress_data + data.rs:79-85 to_bytes + header.rs:60-67.
 * compressor: 0..6 (5 = Auto); error_pct: the CLI's -e (u8); speed: -c */
API int64_t atsc_oracle_compress_stream(const double *samples, uint64_t n, int compressor,
                                        uint32_t error_pct, uint32_t speed, uint8_t *out,
                                        uint64_t cap, int *frame_compressors,
                                        uint64_t frame_cap) {
    double *clean = (double *)malloc(sizeof(double) * (n ? n : 1));
    uint64_t cn = atsc_oracle_clean_data(samples, n, clean);
    uint64_t nchunks = atsc_oracle_chunk_sizes(cn, NULL, 0);
    uint64_t *chunks = (uint64_t *)malloc(sizeof(uint64_t) * (nchunks ? nchunks : 1));
    atsc_oracle_chunk_sizes(cn, chunks, nchunks);
    float max_error = (float)error_pct / 100.0f; /* main.rs:157 `error as f32 / 100.0` */
    buf_t body = {0};
    enc_varint(&body, nchunks);
    uint8_t frame_count = 0;
    uint64_t off = 0;
    int rc = 0;
    for (uint64_t i = 0; i < nchunks && !rc; i++) {
        const double *d = clean + off;
        uint64_t len = chunks[i];
        frame_t f;
        memset(&f, 0, sizeof(f));
        f.sample_count = len;
        f.compressor = compressor;
        int lossy = compressor == C_FFT || compressor == C_POLY || compressor == C_IDW ||
                    compressor == C_AUTO;
        if (lossy) {
            if (compressor == C_AUTO) {
                rc = compress_best(&f, d, len, max_error, speed);
            } else {
                /* frame/mod.rs:65-68 -> Compressor::compress_bounded */
                result_t r;
                rc = get_compress_bounded_results(compressor, d, len, (double)max_error, &r);
                f.data = r.bytes;
            }
        } else {
            rc = compress_unbounded(compressor, d, len, &f.data);
        }
        if (rc) break;
        if (frame_compressors && i < frame_cap) frame_compressors[i] = f.compressor;
        /* frame/mod.rs:25-33 derive(Encode): frame_size (=41), sample_count, compressor, data */
        enc_varint(&body, 41);
        enc_varint(&body, f.sample_count);
        enc_varint(&body, (uint64_t)f.compressor);
        enc_varint(&body, f.data.len);
        buf_put(&body, f.data.p, f.data.len);
        buf_free(&f.data);
        frame_count++; /* header.rs:52-54, u8 wraps in release */
        off += len;
    }
    free(chunks);
    free(clean);
    if (rc) {
        buf_free(&body);
        return -1;
    }
    buf_t all = {0};
    buf_put(&all, "BRRO", 4);
    uint32_t ver = 1;
    buf_put(&all, &ver, 4);
    buf_u8(&all, frame_count);
    buf_put(&all, body.p, body.len);
    buf_free(&body);
    return emit(&all, out, cap);
}

/* data.rs:89-109 from_bytes + decompress.  Returns sample count (or <0).
 * If out==NULL only counts. */
API int64_t atsc_oracle_decompress_stream(const uint8_t *bro, uint64_t len, double *out,
                                          uint64_t cap) {
    if (len < 9) return -1;
    if (memcmp(bro, "BRRO", 4) != 0) return -2;
    uint32_t ver;
    memcpy(&ver, bro + 4, 4);
    if (ver > 1) return -3;
    rd_t r = {bro + 9, (size_t)(len - 9), 0, 0};
    uint64_t nframes = rd_varint(&r);
    uint64_t total = 0;
    for (uint64_t i = 0; i < nframes; i++) {
        (void)rd_varint(&r); /* frame_size */
        uint64_t sc = rd_varint(&r);
        uint64_t comp = rd_varint(&r);
        uint64_t dl = rd_varint(&r);
        if (r.err || r.pos + dl > r.len) return -4;
        const uint8_t *pl = r.p + r.pos;
        r.pos += dl;
        if (comp == C_NOOP) {
            /* noop ignores sample_count: length comes from the payload */
            rd_t q = {pl, (size_t)dl, 0, 0};
            (void)rd_u8(&q);
            uint64_t k = rd_varint(&q);
            if (q.err) return -4;
            if (out) {
                if (total + k > cap) return -5;
                if (noop_to_data(pl, dl, out + total, k) < 0) return -4;
            }
            total += k;
        } else {
            if (out) {
                if (total + sc > cap) return -5;
                if (decompress_payload((int)comp, sc, pl, dl, out + total) < 0) return -4;
            }
            total += sc;
        }
    }
    return (int64_t)total;
}

/* ------------------------------------------------------------------ */
/* threaded batch driver: CPU baseline for bench.py                    */
/* (the reference is single-threaded per series, main.rs:146; one      */
/* worker per series across host cores)                                */
/* ------------------------------------------------------------------ */
typedef struct {
    const double *samples;
    uint64_t series_len, n_series;
    int compressor;
    uint32_t error_pct, speed;
    uint64_t *out_bytes;
    volatile uint64_t *next;
    int fail;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    uint64_t cap = j->series_len * 10 + 4096;
    uint8_t *out = (uint8_t *)malloc(cap);
    for (;;) {
        uint64_t s = __sync_fetch_and_add(j->next, 1);
        if (s >= j->n_series) break;
        int64_t n = atsc_oracle_compress_stream(j->samples + s * j->series_len, j->series_len,
                                                j->compressor, j->error_pct, j->speed, out, cap,
                                                NULL, 0);
        if (n < 0) j->fail = 1;
        j->out_bytes[s] = n < 0 ? 0 : (uint64_t)n;
    }
    free(out);
    return NULL;
}

API int atsc_oracle_compress_batch(const double *samples, uint64_t series_len, uint64_t n_series,
                                   int compressor, uint32_t error_pct, uint32_t speed,
                                   int n_threads, uint64_t *out_bytes) {
    if (n_threads < 1) n_threads = 1;
    volatile uint64_t next = 0;
    batch_job job = {samples, series_len, n_series, compressor, error_pct, speed,
                     out_bytes, &next, 0};
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    for (int i = 0; i < n_threads; i++) pthread_create(&th[i], NULL, batch_worker, &job);
    for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
    free(th);
    return job.fail ? -1 : 0;
}

typedef struct {
    const uint8_t *bros;
    const uint64_t *bro_off; /* n+1 offsets */
    uint64_t n_series;
    double *out;
    uint64_t series_len;
    volatile uint64_t *next;
    int fail;
} dbatch_job;

static void *dbatch_worker(void *arg) {
    dbatch_job *j = (dbatch_job *)arg;
    for (;;) {
        uint64_t s = __sync_fetch_and_add(j->next, 1);
        if (s >= j->n_series) break;
        int64_t n = atsc_oracle_decompress_stream(j->bros + j->bro_off[s],
                                                  j->bro_off[s + 1] - j->bro_off[s],
                                                  j->out + s * j->series_len, j->series_len);
        if (n < 0) j->fail = 1;
    }
    return NULL;
}

API int atsc_oracle_decompress_batch(const uint8_t *bros, const uint64_t *bro_off,
                                     uint64_t n_series, uint64_t series_len, int n_threads,
                                     double *out) {
    if (n_threads < 1) n_threads = 1;
    volatile uint64_t next = 0;
    dbatch_job job = {bros, bro_off, n_series, out, series_len, &next, 0};
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    for (int i = 0; i < n_threads; i++) pthread_create(&th[i], NULL, dbatch_worker, &job);
    for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
    free(th);
    return job.fail ? -1 : 0;
}

// ingest.cpp -- host I/O either side of the hot path:
//   WBRO container   wavbrro/src/wavbrro.rs:24,36-45,78-132, read.rs:23-37, write.rs:21-27
//   CSV value reader atsc/src/csv.rs:36-110
// Host-only C ABI (no GPU needed); used by the `atsc` CLI (atsc_b200/host/atsc_cli.cpp).
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/atsc_gpu.h"

namespace {

constexpr uint32_t WBRO_CHUNK = 2048;  // wavbrro.rs:24 MAX_CHUNK_SIZE

bool parse_f64(const char *b, const char *e, double &out) {
    // Rust `str::parse::<f64>`: no surrounding whitespace, optional sign, inf / infinity / nan
    if (b == e) return false;
    const char *p = b;
    bool neg = false;
    if (*p == '+' || *p == '-') {
        neg = *p == '-';
        p++;
    }
    std::string low(p, e);
    for (auto &c : low) c = (char)tolower(c);
    if (low == "inf" || low == "infinity") {
        out = neg ? -INFINITY : INFINITY;
        return true;
    }
    if (low == "nan") {
        out = NAN;
        return true;
    }
    if (p == e || !((*p >= '0' && *p <= '9') || *p == '.')) return false;
    double v = 0;
    auto r = std::from_chars(p, e, v, std::chars_format::general);
    if (r.ec == std::errc::result_out_of_range) {
        // from_chars leaves v unmodified on range errors; Rust saturates to inf / 0
        char *endp = nullptr;
        std::string tmp(p, e);
        v = strtod(tmp.c_str(), &endp);
        if (endp != tmp.c_str() + tmp.size()) return false;
    } else if (r.ec != std::errc() || r.ptr != e) {
        return false;
    }
    out = neg ? -v : v;
    return true;
}

// one CSV record -> fields (RFC-4180 quoting as the `csv` crate's defaults)
void split_record(const std::string &line, std::vector<std::string> &f) {
    f.clear();
    std::string cur;
    bool q = false;
    for (size_t i = 0; i < line.size(); i++) {
        char c = line[i];
        if (q) {
            if (c == '"' && i + 1 < line.size() && line[i + 1] == '"') {
                cur.push_back('"');
                i++;
            } else if (c == '"')
                q = false;
            else
                cur.push_back(c);
        } else if (c == '"' && cur.empty())
            q = true;
        else if (c == ',') {
            f.push_back(cur);
            cur.clear();
        } else
            cur.push_back(c);
    }
    f.push_back(cur);
}

}  // namespace

extern "C" {

// WavBrro::from_bytes + get_samples (wavbrro.rs:95-132) on a whole .wbro file image.
// Returns the sample count (copies min(count, cap) samples), or <0: -1 bad header, -2 corrupt.
int64_t atsc_wbro_decode(const uint8_t *file, uint64_t len, double *out, uint64_t cap) {
    if (len < 12 || memcmp(file, "WBRO", 4) != 0 || memcmp(file + 8, "WBRO", 4) != 0) return -1;  // read.rs:24-31
    const uint8_t *b = file + 12;
    uint64_t n = len - 12;
    if (n < 16) return -2;
    // rkyv 0.7 archive: the root object is the last 16 bytes
    //   { chunks: (rel_ptr i32, len u32), sample_count u32, bitdepth u8, pad[3] }
    uint64_t root = n - 16;
    int32_t rel;
    uint32_t n_chunks, sample_count;
    memcpy(&rel, b + root, 4);
    memcpy(&n_chunks, b + root + 4, 4);
    memcpy(&sample_count, b + root + 8, 4);
    int64_t arr = (int64_t)root + rel;
    if (arr < 0 || (uint64_t)arr + 8ull * n_chunks > n) return -2;
    uint64_t total = 0;
    for (uint32_t c = 0; c < n_chunks; c++) {
        uint64_t at = (uint64_t)arr + 8ull * c;
        int32_t crel;
        uint32_t clen;
        memcpy(&crel, b + at, 4);
        memcpy(&clen, b + at + 4, 4);
        int64_t data = (int64_t)at + crel;
        if (data < 0 || (uint64_t)data + 8ull * clen > n) return -2;
        for (uint32_t i = 0; i < clen; i++) {
            if (total < cap) memcpy(out + total, b + data + 8ull * i, 8);
            total++;
        }
    }
    (void)sample_count;  // the reference does not cross-check it either
    return (int64_t)total;
}

// WavBrro::to_file_with_data (wavbrro.rs:114-118, write.rs:21-27): 12-byte header + archive.
// Returns the file size (writes only if it fits in cap).
uint64_t atsc_wbro_encode(const double *samples, uint64_t n, uint8_t *out, uint64_t cap) {
    const uint64_t n_chunks = (n + WBRO_CHUNK - 1) / WBRO_CHUNK;
    const uint64_t size = 12 + n * 8 + n_chunks * 8 + 16;
    if (size > cap || !out) return size;
    memcpy(out, "WBRO0000WBRO", 12);
    uint8_t *b = out + 12;
    memcpy(b, samples, n * 8);  // chunk payloads back to back (each 2048 * 8 bytes, 8-aligned)
    uint64_t arr = n * 8;
    for (uint64_t c = 0; c < n_chunks; c++) {
        uint64_t at = arr + 8 * c;
        int32_t rel = (int32_t)((int64_t)(c * WBRO_CHUNK * 8) - (int64_t)at);
        uint32_t clen = (uint32_t)std::min<uint64_t>(WBRO_CHUNK, n - c * WBRO_CHUNK);
        memcpy(b + at, &rel, 4);
        memcpy(b + at + 4, &clen, 4);
    }
    uint64_t root = arr + 8 * n_chunks;
    int32_t rel = (int32_t)((int64_t)arr - (int64_t)root);
    uint32_t nc = (uint32_t)n_chunks, sc = (uint32_t)n;  // `sample_count as u32` truncates (wavbrro.rs:81)
    memcpy(b + root, &rel, 4);
    memcpy(b + root + 4, &nc, 4);
    memcpy(b + root + 8, &sc, 4);
    uint8_t tail[4] = {5, 0, 0, 0};  // bitdepth 5 = f64 (wavbrro.rs:66)
    memcpy(b + root + 12, tail, 4);
    return size;
}

// csv.rs:36-98.  has_header != 0: read_samples_with_headers(time_field, value_field);
// else read_samples (first column).  Returns the value count (copies min(count, cap)), or
// -1 time field missing, -2 value field missing, -3 a value failed to parse, -4 short record.
int64_t atsc_csv_read_values(const char *text, uint64_t len, int has_header, const char *time_field,
                             const char *value_field, double *out, uint64_t cap) {
    std::vector<std::string> f;
    uint64_t pos = 0, count = 0;
    size_t value_idx = 0;
    bool first = true;
    while (pos < len) {
        uint64_t e = pos;
        while (e < len && text[e] != '\n') e++;
        uint64_t le = e;
        if (le > pos && text[le - 1] == '\r') le--;
        std::string line(text + pos, text + le);
        pos = e + 1;
        if (line.empty()) continue;  // the csv crate skips empty lines
        split_record(line, f);
        if (first && has_header) {
            first = false;
            bool tfound = false, vfound = false;
            for (size_t i = 0; i < f.size(); i++) {
                if (!tfound && f[i] == time_field) tfound = true;
                if (!vfound && f[i] == value_field) {
                    vfound = true;
                    value_idx = i;
                }
            }
            if (!tfound) return -1;
            if (!vfound) return -2;
            continue;
        }
        first = false;
        if (value_idx >= f.size()) return -4;
        double v;
        const std::string &s = f[value_idx];
        if (!parse_f64(s.data(), s.data() + s.size(), v)) return -3;
        if (count < cap && out) out[count] = v;
        count++;
    }
    return (int64_t)count;
}

// contiguous frame ranges balanced by sample count: the sharding rule of a multi-device
// context and of bench.py's ranks (frames are independent, frame/mod.rs:71; no collective).
// out_first has n_parts + 1 entries.
void atsc_plan_shards(const uint32_t *frame_len, uint32_t n_frames, uint32_t n_parts, uint32_t *out_first) {
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_frames; i++) total += frame_len[i];
    uint64_t acc = 0;
    uint32_t d = 0;
    out_first[0] = 0;
    for (uint32_t i = 0; i < n_frames; i++) {
        while (d + 1 < n_parts && acc >= (total * (d + 1)) / n_parts) {
            d++;
            out_first[d] = i;
        }
        acc += frame_len[i];
    }
    while (d + 1 <= n_parts) {
        d++;
        out_first[d] = n_frames;
    }
}

}  // extern "C"

// stream.cpp -- host side of the path above the frame ABI: planner, BRO stream layout.
//   main.rs:130-172  compress_data / decompress_data
//   optimizer/mod.rs:47-109  OptimizerPlan (clean_data, get_chunks_sizes, get_execution)
//   data.rs:79-109   CompressedStream::{to_bytes, from_bytes, decompress}
//   header.rs:60-84  CompressorHeader
//   frame/mod.rs:25-33 CompressorFrame (bincode derive: frame_size, sample_count, compressor, data)
// The reference's host language is Rust; no Rust toolchain exists in this image, so this
// layer is C++ over the same C ABI a Rust -sys crate would bind.
#include <algorithm>
#include <cmath>
#include <functional>
#include <cstring>
#include <string>
#include <thread>
#include <memory>
#include <new>
#include <vector>

#include "../../include/atsc_gpu.h"
#include "host_util.h"

using namespace atsc_host;

extern "C" uint64_t atsc_plan_chunk_sizes(uint64_t len, uint32_t *out_sizes, uint64_t cap) {
    std::vector<uint32_t> v;
    chunk_sizes(len, v);
    for (size_t i = 0; i < v.size() && i < cap; i++) out_sizes[i] = v[i];
    return v.size();
}

namespace {

// optimizer/mod.rs:64-71: does the series contain NaN / +-inf?
bool needs_cleaning(const double *p, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) {
        uint64_t b;
        memcpy(&b, p + i, 8);
        if (((b >> 52) & 0x7FF) == 0x7FF) return true;
    }
    return false;
}

void parallel_for(uint32_t n, const std::function<void(uint32_t)> &fn) {
    unsigned hw = std::thread::hardware_concurrency();
    uint32_t nt = std::min<uint32_t>(n, hw ? hw : 4);
    if (nt <= 1) {
        for (uint32_t i = 0; i < n; i++) fn(i);
        return;
    }
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < nt; t++)
        th.emplace_back([&, t]() {
            for (uint32_t i = t; i < n; i += nt) fn(i);
        });
    for (auto &x : th) x.join();
}

}  // namespace

extern "C" int atsc_gpu_compress_series(atsc_ctx *ctx, const double *samples, const uint64_t *series_off,
                                        const uint64_t *series_len, uint32_t n_series, uint8_t compressor,
                                        uint32_t error_pct, uint32_t speed, uint8_t *bro_buf, uint64_t bro_cap,
                                        uint64_t *bro_off, uint64_t *bro_len, uint8_t *frame_near_tie_any) {
    if (!ctx || !samples || !series_off || !series_len || !bro_off || !bro_len) return ATSC_ERR_ARG;
    if (compressor > 6 || error_pct > 50 || speed > 6) return ATSC_ERR_ARG;  // main.rs:187,200
    // ---- clean_data: only series that hold NaN / inf are copied
    std::vector<uint8_t> dirty(n_series, 0);
    parallel_for(n_series, [&](uint32_t s) { dirty[s] = needs_cleaning(samples + series_off[s], series_len[s]); });
    uint64_t extra = 0;
    for (uint32_t s = 0; s < n_series; s++)
        if (dirty[s]) extra += series_len[s];
    // cleaned copies live in one side buffer addressed *relative to `samples`* is impossible, so
    // when anything is dirty the whole batch is staged into one contiguous buffer
    std::vector<double> staged;
    std::vector<uint64_t> c_off(n_series), c_len(n_series);
    const double *base = samples;
    if (extra) {
        uint64_t total = 0;
        for (uint32_t s = 0; s < n_series; s++) total += series_len[s];
        staged.resize(total);
        uint64_t o = 0;
        for (uint32_t s = 0; s < n_series; s++) {
            const double *p = samples + series_off[s];
            uint64_t k = 0;
            if (dirty[s]) {
                for (uint64_t i = 0; i < series_len[s]; i++)
                    if (!(std::isnan(p[i]) || std::isinf(p[i]))) staged[o + k++] = p[i];
            } else {
                memcpy(&staged[o], p, series_len[s] * 8);
                k = series_len[s];
            }
            c_off[s] = o;
            c_len[s] = k;
            o += k;
        }
        base = staged.data();
    } else {
        for (uint32_t s = 0; s < n_series; s++) {
            c_off[s] = series_off[s];
            c_len[s] = series_len[s];
        }
    }
    // ---- get_chunks_sizes / get_execution: one frame list for the whole batch
    std::vector<uint64_t> f_off;
    std::vector<uint32_t> f_len, first_frame(n_series + 1, 0);
    std::vector<uint32_t> cs;
    for (uint32_t s = 0; s < n_series; s++) {
        cs.clear();
        chunk_sizes(c_len[s], cs);
        first_frame[s] = (uint32_t)f_len.size();
        uint64_t o = c_off[s];
        for (uint32_t c : cs) {
            f_off.push_back(o);
            f_len.push_back(c);
            o += c;
        }
    }
    first_frame[n_series] = (uint32_t)f_len.size();
    const uint32_t nf = (uint32_t)f_len.size();
    // main.rs:150-161: lossy compressors take the error bound, the others just compress
    const bool lossy = compressor == ATSC_FFT || compressor == ATSC_POLYNOMIAL || compressor == ATSC_IDW ||
                       compressor == ATSC_AUTO;
    const float max_error = (float)error_pct / 100.0f;  // `arguments.error as f32 / 100.0`
    std::vector<atsc_frame_out> fo(nf);
    // Payload staging: uninitialised, ~2 B per sample to start with (real fleets need about 1); the call reports
    // the size it needed when that is not enough (ATSC_ERR_CAPACITY + payload_used) and is repeated once with it.
    // The worst case is 16 B per sample (RLE of all-distinct f64): sizing -- and zero-filling -- for it up front
    // cost 8 GiB of host memory per 256 Mi-sample batch.
    uint64_t total_len = 0;
    for (uint32_t i = 0; i < nf; i++) total_len += f_len[i];
    uint64_t pcap = total_len * 2 + 64 * (uint64_t)nf + 4096;
    std::unique_ptr<uint8_t[]> payload_mem;
    uint64_t pused = 0;
    if (nf) {
        for (int attempt = 0;; attempt++) {
            payload_mem.reset(new (std::nothrow) uint8_t[pcap]);
            if (!payload_mem) return ATSC_ERR_CAPACITY;
            int rc = atsc_gpu_compress_frames(ctx, base, f_off.data(), f_len.data(), nf, compressor, max_error, speed,
                                              lossy ? 1 : 0, fo.data(), payload_mem.get(), pcap, &pused);
            if (rc == ATSC_ERR_CAPACITY && attempt == 0 && pused > pcap) {
                pcap = pused + 64;
                continue;
            }
            if (rc) return rc;
            break;
        }
    }
    const uint8_t *payload = payload_mem.get();
    // ---- CompressedStream::to_bytes per series
    uint64_t w = 0;
    bool overflow = false;
    std::vector<uint8_t> hdr;
    for (uint32_t s = 0; s < n_series; s++) {
        uint32_t a = first_frame[s], b = first_frame[s + 1];
        uint64_t start = w;
        hdr.clear();
        hdr.insert(hdr.end(), {'B', 'R', 'R', 'O', 1, 0, 0, 0});  // magic + version 1 LE (header.rs:23,60-67)
        hdr.push_back((uint8_t)((b - a) & 0xFF));                    // frame_count: u8 += 1 per frame (header.rs:52)
        put_varint(hdr, b - a);                                      // Vec<CompressorFrame> length
        auto emit = [&](const uint8_t *p, uint64_t n) {
            if (w + n > bro_cap)
                overflow = true;
            else if (bro_buf)
                memcpy(bro_buf + w, p, n);
            w += n;
        };
        emit(hdr.data(), hdr.size());
        uint8_t tie = 0;
        for (uint32_t i = a; i < b; i++) {
            hdr.clear();
            put_varint(hdr, 41);  // frame_size: size_of_val sum, always 41 on 64-bit (frame/mod.rs:50-56)
            put_varint(hdr, f_len[i]);
            put_varint(hdr, fo[i].compressor);
            put_varint(hdr, fo[i].payload_len);
            emit(hdr.data(), hdr.size());
            emit(payload + fo[i].payload_off, fo[i].payload_len);
            tie |= fo[i].near_tie;
        }
        bro_off[s] = start;
        bro_len[s] = w - start;
        if (frame_near_tie_any) frame_near_tie_any[s] = tie;
    }
    return overflow ? ATSC_ERR_CAPACITY : ATSC_OK;
}

extern "C" int atsc_gpu_decompress_series(atsc_ctx *ctx, const uint8_t *bro_buf, const uint64_t *bro_off,
                                          const uint64_t *bro_len, uint32_t n_series, double *out_samples,
                                          const uint64_t *out_off, uint64_t *out_count) {
    if (!ctx || !bro_buf || !bro_off || !bro_len) return ATSC_ERR_ARG;
    std::vector<atsc_frame_in> frames;
    uint64_t max_end = 0;
    for (uint32_t s = 0; s < n_series; s++) {
        const uint8_t *p = bro_buf + bro_off[s];
        const uint64_t len = bro_len[s];
        if (len < 9) return ATSC_ERR_FORMAT;                    // data.rs:93 split_at(9)
        if (memcmp(p, "BRRO", 4) != 0) return ATSC_ERR_FORMAT;  // header.rs:72 "Magic bytes are not correct!"
        uint32_t ver;
        memcpy(&ver, p + 4, 4);
        if (ver > 1) return ATSC_ERR_FORMAT;  // header.rs:30-37 newer file version
        uint64_t pos = 9, nfr;
        if (!get_varint(p, len, pos, nfr)) return ATSC_ERR_FORMAT;
        uint64_t total = 0;
        for (uint64_t i = 0; i < nfr; i++) {
            uint64_t fsz, sc, comp, dl;
            if (!get_varint(p, len, pos, fsz) || !get_varint(p, len, pos, sc) || !get_varint(p, len, pos, comp) ||
                !get_varint(p, len, pos, dl))
                return ATSC_ERR_FORMAT;
            if (dl > len - pos || comp > 6) return ATSC_ERR_FORMAT;  // (pos <= len here; no wrap-around)
            if (comp == ATSC_NOOP) {
                // noop_to_data ignores sample_count: the stored Vec<i64> decides (noop.rs:79-83)
                uint64_t q = pos + 1, k;
                if (dl < 2 || !get_varint(p, pos + dl, q, k)) return ATSC_ERR_FORMAT;
                sc = k;
            }
            if (sc > 131072) return ATSC_ERR_FORMAT;  // no writer produces a frame beyond MAX_FRAME_SIZE (optimizer/mod.rs:25)
            if (out_samples && sc) {
                atsc_frame_in f;
                memset(&f, 0, sizeof f);
                f.compressor = (uint8_t)comp;
                f.sample_count = (uint32_t)sc;
                f.payload_off = bro_off[s] + pos;
                f.payload_len = (uint32_t)dl;
                f.out_off = (out_off ? out_off[s] : 0) + total;
                frames.push_back(f);
            }
            max_end = std::max(max_end, bro_off[s] + pos + dl);
            pos += dl;
            total += sc;
        }
        if (out_count) out_count[s] = total;
    }
    if (!out_samples || frames.empty()) return ATSC_OK;
    return atsc_gpu_decompress_frames(ctx, frames.data(), (uint32_t)frames.size(), bro_buf, max_end, out_samples);
}

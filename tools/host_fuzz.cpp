// host_fuzz.cpp -- mutation fuzzing of the host-side parsers (WBRO, CSV, VSRI text, BRO stream layout)
// under AddressSanitizer / UBSan.  Build and run (no GPU needed; the two device entry points the
// stream layer calls are stubbed):
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=all tools/host_fuzz.cpp \
//       atsc_b200/csrc/ingest.cpp atsc_b200/csrc/vsri.cpp atsc_b200/csrc/stream.cpp -o /tmp/host_fuzz && /tmp/host_fuzz 200000
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../include/atsc_gpu.h"

// ---- stubs: what the stream layer hands to the device is checked here instead of executed
static uint64_t g_payload_bytes = 0;
extern "C" int atsc_gpu_compress_frames(atsc_ctx *, const double *, const uint64_t *, const uint32_t *, uint32_t, uint8_t, float,
                                        uint32_t, int, atsc_frame_out *, uint8_t *, uint64_t, uint64_t *) {
    return ATSC_ERR_CUDA;
}
extern "C" int atsc_gpu_decompress_frames(atsc_ctx *, const atsc_frame_in *f, uint32_t n, const uint8_t *, uint64_t payload_bytes,
                                          double *) {
    for (uint32_t i = 0; i < n; i++) {
        if (f[i].payload_off + f[i].payload_len > payload_bytes || payload_bytes > g_payload_bytes || f[i].sample_count > 131072 ||
            f[i].out_off + f[i].sample_count > (1u << 20)) {
            fprintf(stderr, "stream layer passed an out-of-range frame to the device\n");
            abort();
        }
    }
    return ATSC_OK;
}

// a small valid BRO stream: header + two frames (Constant, Noop)
static std::string make_bro() {
    std::string b = "BRRO";
    b += std::string("\x01\x00\x00\x00", 4);
    b += '\x02';                  // frame_count u8
    b += '\x02';                  // varint n_frames
    b += std::string("\x29\xfb\x00\x04\x03\x03\x1e\x03\x01", 9);             // 41, 1024, Constant, len 3, [30, 3, 1]
    b += std::string("\x29\x05\x00\x07\xfa\x05\x02\x02\x02\x02\x02", 11);   // 41, 5, Noop, len 7, [250, 5, 2 x5]
    return b;
}

int main(int argc, char **argv) {
    const long iters = argc > 1 ? atol(argv[1]) : 100000;
    std::mt19937_64 rng(12345);
    // seeds: a valid WBRO file, a CSV, an index
    std::vector<double> v(700);
    for (size_t i = 0; i < v.size(); i++) v[i] = (double)(i % 17) * 0.5 - 3.0;
    std::vector<uint8_t> wbro(atsc_wbro_encode(v.data(), v.size(), nullptr, 0));
    atsc_wbro_encode(v.data(), v.size(), wbro.data(), wbro.size());
    const std::string csv = "time,value,extra\n1,1.5,a\n2,-2.25e3,b\n3,nan,c\n\"4\",\"7\",d\n";
    const std::string idx = "55745\n59435\n15,0,55745,166\n15,166,58505,63\n";
    const std::string bro = make_bro();
    std::vector<double> out(4096);
    std::vector<double> big(1u << 20);
    long parsed = 0;
    {  // the unmutated seeds must parse
        uint64_t off = 0, len = bro.size(), count = 0;
        g_payload_bytes = bro.size();
        atsc_vsri *x = atsc_vsri_from_text(idx.data(), idx.size());
        if (atsc_wbro_decode(wbro.data(), wbro.size(), nullptr, 0) != 700 ||
            atsc_csv_read_values(csv.data(), csv.size(), 1, "time", "value", nullptr, 0) < 0 || !x ||
            atsc_gpu_decompress_series(reinterpret_cast<atsc_ctx *>(&off), (const uint8_t *)bro.data(), &off, &len, 1, nullptr, nullptr,
                                       &count) != ATSC_OK ||
            count != 1029) {
            fprintf(stderr, "a seed does not parse (count %llu)\n", (unsigned long long)count);
            return 1;
        }
        atsc_vsri_free(x);
    }
    for (long it = 0; it < iters; it++) {
        const int which = (int)(rng() % 4);
        std::string buf = which == 0 ? std::string((const char *)wbro.data(), wbro.size()) : which == 1 ? csv : which == 2 ? idx : bro;
        const int muts = 1 + (int)(rng() % 4);
        for (int m = 0; m < muts && !buf.empty(); m++) {
            const size_t pos = rng() % buf.size();
            switch (rng() % 5) {
                case 0: buf[pos] = (char)(rng() & 0xFF); break;
                case 1: buf.erase(pos, 1 + rng() % 8); break;
                case 2: buf.insert(pos, 1 + rng() % 4, (char)(rng() & 0xFF)); break;
                case 3: buf.resize(pos); break;
                default: buf[pos] = "0123456789,.-\n\"e"[rng() % 16]; break;
            }
        }
        if (which == 0) {
            int64_t n = atsc_wbro_decode((const uint8_t *)buf.data(), buf.size(), nullptr, 0);
            if (n > 0 && (size_t)n <= out.size()) {
                atsc_wbro_decode((const uint8_t *)buf.data(), buf.size(), out.data(), out.size());
                parsed++;
            }
        } else if (which == 1) {
            int64_t n = atsc_csv_read_values(buf.data(), buf.size(), 1, "time", "value", nullptr, 0);
            if (n > 0 && (size_t)n <= out.size()) {
                atsc_csv_read_values(buf.data(), buf.size(), 1, "time", "value", out.data(), out.size());
                parsed++;
            }
            atsc_csv_read_values(buf.data(), buf.size(), 0, "time", "value", out.data(), 2);  // short buffer
        } else if (which == 3) {
            uint64_t off = 0, len = buf.size(), count = 0, ooff = 0;
            atsc_ctx *fake = reinterpret_cast<atsc_ctx *>(&off);  // never dereferenced by the stream layer
            g_payload_bytes = buf.size();
            int rc = atsc_gpu_decompress_series(fake, (const uint8_t *)buf.data(), &off, &len, 1, nullptr, nullptr, &count);
            if (rc == ATSC_OK && count <= big.size()) {
                rc = atsc_gpu_decompress_series(fake, (const uint8_t *)buf.data(), &off, &len, 1, big.data(), &ooff, &count);
                if (rc == ATSC_OK) parsed++;
            }
        } else {
            atsc_vsri *x = atsc_vsri_from_text(buf.data(), buf.size());
            if (x) {
                parsed++;
                int32_t o = 0;
                for (int32_t q = -2; q < 300; q += 7) {
                    atsc_vsri_get_time(x, q, &o);
                    atsc_vsri_get_sample(x, 55745 + q * 13, &o);
                    atsc_vsri_get_next_sample(x, 55745 + q * 13, &o);
                    atsc_vsri_get_previous_sample(x, 55745 + q * 13, &o);
                    atsc_vsri_is_empty(x, 55000 + q, 56000 + q);
                }
                std::vector<int32_t> ts(64);
                atsc_vsri_all_timestamps(x, ts.data(), ts.size());
                std::string text(atsc_vsri_to_text(x, nullptr, 0), '\0');
                atsc_vsri_to_text(x, text.data(), text.size());
                atsc_vsri_update_for_point(x, atsc_vsri_max(x) + 15);
                atsc_vsri_free(x);
            }
        }
    }
    printf("host_fuzz: %ld inputs, %ld parsed, no sanitizer report\n", iters, parsed);
    return 0;
}

// atsc_cli.cpp -- the `atsc` command line of the reference (atsc/src/main.rs:29-243) on top of
// libatsc_gpu.so.  Same option surface:
//   atsc [--compressor auto|noop|fft|constant|polynomial|idw|rle] [-e|--error 0..50] [-u]
//        [-c|--compression-selection-sample-level 0..6] [--verbose] [--csv] [--no-header]
//        [--fields TIME,VALUE] <INPUT file or directory>
// The reference's host language is Rust (no toolchain in this image): this is the C++ host.
#include <dirent.h>
#include <sys/stat.h>

#include <charconv>
#include <cmath>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "../../include/atsc_gpu.h"
#include "cli_common.h"

namespace {

struct Args {
    std::string input;
    int compressor = ATSC_AUTO;  // default_value = "auto" (main.rs:180)
    unsigned error = 3;          // default_value_t = 3 (main.rs:187)
    bool uncompress = false;
    unsigned speed = 0;
    bool verbose = false, csv = false, no_header = false;
    std::string fields = "time,value";
};

// Rust `{:?}` of f64: shortest round-trip digits, always a fractional part or an exponent
std::string debug_f64(double v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    char buf[64];
    double a = std::fabs(v);
    bool sci = a != 0.0 && (a < 1e-4 || a >= 1e16);
    auto r = std::to_chars(buf, buf + sizeof buf, v, sci ? std::chars_format::scientific : std::chars_format::fixed);
    std::string s(buf, r.ptr);
    if (sci) {
        // 1e+21 -> 1e21, 1.5e-07 -> 1.5e-7
        size_t e = s.find('e');
        std::string mant = s.substr(0, e), ex = s.substr(e + 1);
        bool neg = ex[0] == '-';
        ex = ex.substr(1);
        while (ex.size() > 1 && ex[0] == '0') ex.erase(0, 1);
        return mant + "e" + (neg ? "-" : "") + ex;
    }
    if (s.find('.') == std::string::npos) s += ".0";
    return s;
}
void print_vec(const char *name, const std::vector<double> &v) {
    std::string s = std::string(name) + "=[";
    for (size_t i = 0; i < v.size(); i++) {
        if (i) s += ", ";
        s += debug_f64(v[i]);
    }
    s += "]";
    puts(s.c_str());
}

// compress_data (main.rs:130-166) for one series
int compress_series(atsc_ctx *ctx, const std::vector<double> &data, const Args &a, std::vector<uint8_t> &bro) {
    uint64_t off = 0, len = data.size(), boff = 0, blen = 0;
    bro.resize(data.size() * 16 + 4096 + 64 * (data.size() / 512 + 8));
    double dummy = 0.0;
    int rc = atsc_gpu_compress_series(ctx, data.empty() ? &dummy : data.data(), &off, &len, 1, (uint8_t)a.compressor,
                                      a.error, a.speed, bro.data(), bro.size(), &boff, &blen, nullptr);
    if (rc) return rc;
    bro.resize((size_t)blen);
    return 0;
}

bool load_series(const std::string &path, const Args &a, std::vector<double> &data);

int process_single_file(atsc_ctx *ctx, const std::string &path, const Args &a) {
    std::vector<uint8_t> file;
    if (a.uncompress) {
        // bro_reader::read_file (utils/readers/bro_reader.rs:31-46): needs >= 12 bytes starting "BRRO"
        if (!read_file(path, file)) return 1;
        if (file.size() < 12) {
            fprintf(stderr, "failed to fill whole buffer File: %s\n", path.c_str());
            return 1;
        }
        if (memcmp(file.data(), "BRRO", 4) != 0) return 0;  // not a BRO file: silently skipped
        uint64_t off = 0, len = file.size(), count = 0, ooff = 0;
        int rc = atsc_gpu_decompress_series(ctx, file.data(), &off, &len, 1, nullptr, nullptr, &count);
        if (rc) {
            fprintf(stderr, "decompress failed (%d): %s\n", rc, atsc_gpu_last_error(ctx));
            return 1;
        }
        std::vector<double> out((size_t)count + 1);
        if (count) {
            rc = atsc_gpu_decompress_series(ctx, file.data(), &off, &len, 1, out.data(), &ooff, &count);
            if (rc) {
                fprintf(stderr, "decompress failed (%d): %s\n", rc, atsc_gpu_last_error(ctx));
                return 1;
            }
        }
        out.resize((size_t)count);
        if (a.verbose) print_vec("Output", out);
        std::vector<uint8_t> w(atsc_wbro_encode(out.data(), out.size(), nullptr, 0));
        atsc_wbro_encode(out.data(), out.size(), w.data(), w.size());
        return write_file(with_extension(path, "wbro"), w.data(), w.size()) ? 0 : 1;
    }
    std::vector<double> data;
    if (!load_series(path, a, data)) return 1;
    if (a.verbose) print_vec("Input", data);
    std::vector<uint8_t> bro;
    int rc = compress_series(ctx, data, a, bro);
    if (rc) {
        fprintf(stderr, "compress failed (%d): %s\n", rc, atsc_gpu_last_error(ctx));
        return 1;
    }
    return write_file(with_extension(path, "bro"), bro.data(), bro.size()) ? 0 : 1;
}

// reads one input file of a compression run into `data`; returns false (after reporting) on failure
bool load_series(const std::string &path, const Args &a, std::vector<double> &data) {
    std::vector<uint8_t> file;
    if (!read_file(path, file)) {
        fprintf(stderr, "cannot open %s\n", path.c_str());
        return false;
    }
    if (a.csv) {
        std::string tf = "time", vf = "value";
        size_t comma = a.fields.find(',');
        if (comma != std::string::npos) {
            tf = a.fields.substr(0, comma);
            vf = a.fields.substr(comma + 1);
        }
        int64_t n = atsc_csv_read_values((const char *)file.data(), file.size(), a.no_header ? 0 : 1, tf.c_str(), vf.c_str(),
                                         nullptr, 0);
        if (n < 0) {
            const char *msg = n == -1 ? "Timestamp field is not found" : n == -2 ? "Value field is not found" : "Parsing value is failed";
            fprintf(stderr, "%s File: %s\n", msg, path.c_str());
            return false;
        }
        data.resize((size_t)n);
        atsc_csv_read_values((const char *)file.data(), file.size(), a.no_header ? 0 : 1, tf.c_str(), vf.c_str(), data.data(),
                             data.size());
    } else {
        int64_t n = atsc_wbro_decode(file.data(), file.size(), nullptr, 0);
        if (n < 0) {
            fprintf(stderr, "Ill-formed WAVBRRO file File: %s\n", path.c_str());
            return false;
        }
        data.resize((size_t)n);
        atsc_wbro_decode(file.data(), file.size(), data.data(), data.size());
    }
    return true;
}

// Fleet driver (main.rs:50-68).  The reference walks read_dir lazily, processes every
// file twice and picks up the files it writes on the way; here the listing is taken once and the
// files go to the GPU in BATCHES (up to ~256 Mi samples or 4096 files per call), so the wave
// pipeline of the library sees whole fleets instead of one series at a time.
int process_directory(atsc_ctx *ctx, const Args &a) {
    DIR *d = opendir(a.input.c_str());
    if (!d) return 1;
    std::vector<std::string> files;
    while (dirent *e = readdir(d)) {
        std::string p = a.input + "/" + e->d_name;
        struct stat st;
        if (stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode)) files.push_back(p);
    }
    closedir(d);
    std::sort(files.begin(), files.end());
    int ret = 0;
    const uint64_t BATCH_SAMPLES = 256ull << 20;
    const size_t BATCH_FILES = 4096;
    // One GPU call expands the series [s0, s1) of a batch.  When the call fails (one malformed or truncated
    // .bro fails the whole call) the batch falls back to one call per file, so every healthy file is still
    // written and the offending path is named -- the reference works per file and carries on (main.rs:50-68).
    auto expand = [&](const std::vector<uint8_t> &blob, const std::vector<uint64_t> &off, const std::vector<uint64_t> &len,
                      const std::vector<std::string> &names, uint32_t s0, uint32_t s1, auto &&self) -> void {
        const uint32_t n = s1 - s0;
        std::vector<uint64_t> count(n), ooff(n);
        int rc = atsc_gpu_decompress_series(ctx, blob.data(), off.data() + s0, len.data() + s0, n, nullptr, nullptr, count.data());
        uint64_t total = 0;
        for (uint32_t s = 0; s < n; s++) {
            ooff[s] = total;
            total += count[s];
        }
        std::vector<double> out;
        if (!rc) {
            // a hostile header can claim 131072 samples per 8-byte Constant frame: refuse absurd totals, survive bad_alloc
            uint64_t bytes_in = 0;
            for (uint32_t s = 0; s < n; s++) bytes_in += len[s0 + s];
            if (total > (1ull << 34) || total / 131072 > bytes_in) {
                rc = ATSC_ERR_FORMAT;
            } else {
                try {
                    out.resize((size_t)total + 1);
                } catch (const std::bad_alloc &) {
                    rc = ATSC_ERR_CAPACITY;
                }
            }
        }
        if (!rc && total)
            rc = atsc_gpu_decompress_series(ctx, blob.data(), off.data() + s0, len.data() + s0, n, out.data(), ooff.data(), count.data());
        if (rc) {
            if (n == 1) {
                fprintf(stderr, "decompress failed (%d): %s File: %s\n", rc, atsc_gpu_last_error(ctx), names[s0].c_str());
                ret = 1;
            } else {
                for (uint32_t s = s0; s < s1; s++) self(blob, off, len, names, s, s + 1, self);
            }
            return;
        }
        for (uint32_t s = 0; s < n; s++) {
            const double *p = out.data() + ooff[s];
            if (a.verbose) print_vec("Output", std::vector<double>(p, p + count[s]));
            std::vector<uint8_t> w(atsc_wbro_encode(p, count[s], nullptr, 0));
            atsc_wbro_encode(p, count[s], w.data(), w.size());
            if (!write_file(with_extension(names[s0 + s], "wbro"), w.data(), w.size())) ret = 1;
        }
    };
    if (a.uncompress) {
        size_t i = 0;
        while (i < files.size()) {
            std::vector<uint8_t> blob;
            std::vector<uint64_t> off, len;
            std::vector<std::string> names;
            while (i < files.size() && names.size() < BATCH_FILES && blob.size() < (1ull << 30)) {
                std::vector<uint8_t> file;
                const std::string &path = files[i++];
                if (!read_file(path, file)) {
                    ret = 1;
                    continue;
                }
                if (file.size() < 12) {  // bro_reader.rs:41-43
                    fprintf(stderr, "failed to fill whole buffer File: %s\n", path.c_str());
                    ret = 1;
                    continue;
                }
                if (memcmp(file.data(), "BRRO", 4) != 0) continue;  // not a BRO file: silently skipped
                off.push_back(blob.size());
                len.push_back(file.size());
                names.push_back(path);
                blob.insert(blob.end(), file.begin(), file.end());
            }
            if (!names.empty()) expand(blob, off, len, names, 0, (uint32_t)names.size(), expand);
        }
        return ret;
    }
    // the same for compression: series [s0, s1) of a batch in one call, per file when that fails.  The
    // output buffer starts at ~2 B/sample and grows on ATSC_ERR_CAPACITY (worst case 16 B/sample).
    auto squeeze = [&](const std::vector<double> &samples, const std::vector<uint64_t> &off, const std::vector<uint64_t> &len,
                       const std::vector<std::string> &names, uint32_t s0, uint32_t s1, auto &&self) -> void {
        const uint32_t n = s1 - s0;
        uint64_t ns = 0;
        for (uint32_t s = s0; s < s1; s++) ns += len[s];
        std::vector<uint64_t> boff(n), blen(n);
        std::unique_ptr<uint8_t[]> bro;
        const uint64_t worst = ns * 16 + 4096 * (uint64_t)n + 64 * (ns / 512 + 8 * (uint64_t)n);
        uint64_t cap = std::min<uint64_t>(worst, ns * 2 + 4096 * (uint64_t)n + (1u << 20));
        double dummy = 0.0;
        int rc;
        for (;;) {
            bro.reset(new (std::nothrow) uint8_t[cap]);
            if (!bro) {
                rc = ATSC_ERR_CAPACITY;
                break;
            }
            rc = atsc_gpu_compress_series(ctx, samples.empty() ? &dummy : samples.data(), off.data() + s0, len.data() + s0, n,
                                          (uint8_t)a.compressor, a.error, a.speed, bro.get(), cap, boff.data(), blen.data(), nullptr);
            if (rc != ATSC_ERR_CAPACITY || cap >= worst) break;
            cap = std::min<uint64_t>(worst, cap * 4);
        }
        if (rc) {
            if (n == 1) {
                fprintf(stderr, "compress failed (%d): %s File: %s\n", rc, atsc_gpu_last_error(ctx), names[s0].c_str());
                ret = 1;
            } else {
                for (uint32_t s = s0; s < s1; s++) self(samples, off, len, names, s, s + 1, self);
            }
            return;
        }
        for (uint32_t s = 0; s < n; s++)
            if (!write_file(with_extension(names[s0 + s], "bro"), bro.get() + boff[s], (size_t)blen[s])) ret = 1;
    };
    size_t i = 0;
    while (i < files.size()) {
        std::vector<double> samples;
        std::vector<uint64_t> off, len;
        std::vector<std::string> names;
        while (i < files.size() && names.size() < BATCH_FILES && samples.size() < BATCH_SAMPLES) {
            std::vector<double> data;
            const std::string &path = files[i++];
            if (!load_series(path, a, data)) {
                ret = 1;
                continue;
            }
            if (a.verbose) print_vec("Input", data);
            off.push_back(samples.size());
            len.push_back(data.size());
            names.push_back(path);
            samples.insert(samples.end(), data.begin(), data.end());
        }
        if (!names.empty()) squeeze(samples, off, len, names, 0, (uint32_t)names.size(), squeeze);
    }
    return ret;
}

int usage() {
    fputs("A Time-Series compressor\n\nUsage: atsc [OPTIONS] <INPUT>\n\nOptions:\n"
          "      --compressor <COMPRESSOR>  [default: auto] [possible values: auto, noop, fft, constant, polynomial, idw, rle]\n"
          "  -e, --error <ERROR>            maximum allowed error in percent, 0..50 [default: 3]\n"
          "  -u                             Uncompresses the input file/directory\n"
          "  -c, --compression-selection-sample-level <N>  0..6 [default: 0]\n"
          "      --verbose                  dumps every sample\n"
          "      --csv                      input is a CSV file\n"
          "      --no-header                the CSV has no header\n"
          "      --fields <FIELDS>          TIME_FIELD_NAME,VALUE_FIELD_NAME [default: time,value]\n",
          stderr);
    return 2;
}

}  // namespace

int main(int argc, char **argv) {
    Args a;
    for (int i = 1; i < argc; i++) {
        std::string s = argv[i];
        auto value = [&](const char *name) -> std::string {
            std::string pre = std::string(name) + "=";
            if (s.rfind(pre, 0) == 0) return s.substr(pre.size());
            if (i + 1 < argc) return argv[++i];
            exit(usage());
        };
        if (s == "--compressor" || s.rfind("--compressor=", 0) == 0) {
            std::string v = value("--compressor");
            static const char *names[] = {"noop", "fft", "idw", "constant", "polynomial", "auto", "rle"};
            int c = -1;
            for (int k = 0; k < 7; k++)
                if (v == names[k]) c = k;
            if (c < 0) return usage();
            a.compressor = c;
        } else if (s == "-e" || s == "--error" || s.rfind("--error=", 0) == 0) {
            a.error = (unsigned)atoi(value("--error").c_str());
            if (a.error > 50) return usage();
        } else if (s == "-u")
            a.uncompress = true;
        else if (s == "-c" || s == "--compression-selection-sample-level" ||
                 s.rfind("--compression-selection-sample-level=", 0) == 0) {
            a.speed = (unsigned)atoi(value("--compression-selection-sample-level").c_str());
            if (a.speed > 6) return usage();
        } else if (s == "--verbose")
            a.verbose = true;
        else if (s == "--csv")
            a.csv = true;
        else if (s == "--no-header")
            a.no_header = true;
        else if (s == "--fields" || s.rfind("--fields=", 0) == 0)
            a.fields = value("--fields");
        else if (s == "-h" || s == "--help")
            return usage();
        else if (!s.empty() && s[0] == '-')
            return usage();
        else
            a.input = s;
    }
    if (a.input.empty()) return usage();
    struct stat st;
    if (stat(a.input.c_str(), &st) != 0) {
        fprintf(stderr, "No such file or directory: %s\n", a.input.c_str());
        return 1;
    }
    atsc_ctx *ctx = nullptr;
    int rc = atsc_gpu_create(nullptr, 0, &ctx);
    if (rc) {
        fprintf(stderr, "atsc: no usable CUDA device (status %d); this build has no CPU path\n", rc);
        return 1;
    }
    int ret;
    if (S_ISREG(st.st_mode))
        ret = process_single_file(ctx, a.input, a);
    else if (S_ISDIR(st.st_mode))
        ret = process_directory(ctx, a);
    else {
        fputs("The provided path is neither a file nor a directory.\n", stderr);
        ret = 1;
    }
    atsc_gpu_destroy(ctx);
    return ret;
}

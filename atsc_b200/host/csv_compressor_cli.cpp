// csv_compressor_cli.cpp -- the reference's `csv-compressor` tool (csv-compressor/src/main.rs:32-238,
// csv.rs, metric.rs) on top of libatsc_gpu.so: a `timestamp,value` CSV becomes a BRO stream
// (compressed on the GPU) plus a VSRI timestamp index, and back.  Same option surface:
//   csv-compressor [-o|--output PATH] [-u] [--no-compression] [--output-vsri] [--output-wavbrro]
//                  [--output-csv] [--compressor auto|noop|fft|constant|polynomial|idw]
//                  [-e|--error 0..50] [-c|--compression-selection-sample-level 0..6] <INPUT>
// The reference's host language is Rust (no toolchain in this image): this is the C++ host.
#include <sys/stat.h>

#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/atsc_gpu.h"
#include "cli_common.h"

namespace {

struct Args {
    std::string input, output;
    bool uncompress = false, no_compression = false, output_vsri = false, output_wavbrro = false, output_csv = false;
    int compressor = ATSC_AUTO;  // default_value = "auto" (main.rs:68)
    unsigned error = 5;          // default_value_t = 5 (main.rs:75)
    unsigned speed = 0;
};

struct Sample {  // csv.rs:27-30
    int64_t timestamp;
    double value;
};

// one CSV record -> fields; double quotes group, "" inside quotes is a literal quote
void split_record(const std::string &line, std::vector<std::string> &out) {
    out.clear();
    std::string cur;
    bool quoted = false;
    for (size_t i = 0; i < line.size(); i++) {
        char c = line[i];
        if (quoted) {
            if (c == '"' && i + 1 < line.size() && line[i + 1] == '"') {
                cur += '"';
                i++;
            } else if (c == '"')
                quoted = false;
            else
                cur += c;
        } else if (c == '"' && cur.empty())
            quoted = true;
        else if (c == ',') {
            out.push_back(cur);
            cur.clear();
        } else
            cur += c;
    }
    out.push_back(cur);
}

// csv::read_samples_from_csv_file (csv.rs:43-47): header row names the columns; `timestamp` must
// parse as i64 and `value` as f64, any record that does not is an error for the whole file
bool read_samples(const std::string &path, std::vector<Sample> &out, std::string &err) {
    std::vector<uint8_t> file;
    if (!read_file(path, file)) {
        err = "cannot open " + path;
        return false;
    }
    std::vector<std::string> f;
    int tcol = -1, vcol = -1;
    bool header = true;
    size_t pos = 0, lineno = 0;
    const std::string text((const char *)file.data(), file.size());
    while (pos < text.size()) {
        size_t nl = text.find('\n', pos);
        if (nl == std::string::npos) nl = text.size();
        std::string line = text.substr(pos, nl - pos);
        pos = nl + 1;
        lineno++;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) continue;  // the csv crate skips empty lines
        split_record(line, f);
        if (header) {
            for (size_t k = 0; k < f.size(); k++) {
                if (f[k] == "timestamp" && tcol < 0) tcol = (int)k;
                if (f[k] == "value" && vcol < 0) vcol = (int)k;
            }
            if (tcol < 0 || vcol < 0) {
                err = std::string("missing field `") + (tcol < 0 ? "timestamp" : "value") + "`";
                return false;
            }
            header = false;
            continue;
        }
        if ((int)f.size() <= tcol || (int)f.size() <= vcol) {
            err = "record on line " + std::to_string(lineno) + " has too few fields";
            return false;
        }
        Sample s;
        const std::string &ts = f[(size_t)tcol], &vs = f[(size_t)vcol];
        const char *tb = ts.c_str() + (ts.size() > 1 && ts[0] == '+' ? 1 : 0);
        auto r = std::from_chars(tb, ts.c_str() + ts.size(), s.timestamp);
        if (r.ec != std::errc() || r.ptr != ts.c_str() + ts.size() || ts.empty()) {
            err = "line " + std::to_string(lineno) + ": invalid timestamp `" + ts + "`";
            return false;
        }
        // str::parse::<f64>: decimal/exponent forms, "inf", "infinity", "nan" (any case), optional sign
        const char *vb = vs.c_str() + (vs.size() > 1 && vs[0] == '+' ? 1 : 0);
        auto q = std::from_chars(vb, vs.c_str() + vs.size(), s.value);
        if (q.ec != std::errc() || q.ptr != vs.c_str() + vs.size() || vs.empty()) {
            err = "line " + std::to_string(lineno) + ": invalid value `" + vs + "`";
            return false;
        }
        out.push_back(s);
    }
    return true;
}

// what the csv crate's serializer writes for an f64 (ryu): shortest round-trip digits, plain
// decimals for 1e-5 <= |v| < 1e16, `d.ddde±x` outside, always with a fractional part when plain
std::string ryu_f64(double v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    char buf[64];
    const double a = std::fabs(v);
    const bool sci = a != 0.0 && (a < 1e-5 || a >= 1e16);
    auto r = std::to_chars(buf, buf + sizeof buf, v, sci ? std::chars_format::scientific : std::chars_format::fixed);
    std::string s(buf, r.ptr);
    if (sci) {
        size_t e = s.find('e');
        std::string mant = s.substr(0, e), ex = s.substr(e + 1);
        const bool neg = ex[0] == '-';
        ex = ex.substr(1);
        while (ex.size() > 1 && ex[0] == '0') ex.erase(0, 1);
        return mant + "e" + (neg ? "-" : "") + ex;
    }
    if (s.find('.') == std::string::npos) s += ".0";
    return s;
}

// csv::write_samples_to_csv_file (csv.rs:50-58)
bool write_samples(const std::string &path, const std::vector<Sample> &samples) {
    std::string s = samples.empty() ? "" : "timestamp,value\n";
    for (const Sample &x : samples) s += std::to_string(x.timestamp) + "," + ryu_f64(x.value) + "\n";
    return write_file(path, (const uint8_t *)s.data(), s.size());
}

bool write_wbro(const std::string &path, const std::vector<double> &v) {
    std::vector<uint8_t> w(atsc_wbro_encode(v.data(), v.size(), nullptr, 0));
    atsc_wbro_encode(v.data(), v.size(), w.data(), w.size());
    return write_file(path, w.data(), w.size());
}

int fail(const std::string &msg) {
    fprintf(stderr, "csv-compressor: %s\n", msg.c_str());
    return 101;  // the reference panics (expect/unwrap) on every one of these: exit status 101
}

// process_args, uncompress branch (main.rs:148-186)
int uncompress(atsc_ctx *ctx, const Args &a, const std::string &output_base) {
    std::vector<uint8_t> file;
    // bro_reader::read_file (atsc/src/utils/readers/bro_reader.rs:31-46)
    if (!read_file(a.input, file) || file.size() < 12) return fail("failed to read bro file");
    if (memcmp(file.data(), "BRRO", 4) != 0) return 0;  // Ok(None): nothing to do
    uint64_t off = 0, len = file.size(), count = 0, ooff = 0;
    int rc = atsc_gpu_decompress_series(ctx, file.data(), &off, &len, 1, nullptr, nullptr, &count);
    std::vector<double> data((size_t)count + 1);
    if (!rc && count) rc = atsc_gpu_decompress_series(ctx, file.data(), &off, &len, 1, data.data(), &ooff, &count);
    if (rc) return fail(std::string("decompress failed: ") + atsc_gpu_last_error(ctx));
    data.resize((size_t)count);

    std::vector<uint8_t> text;
    if (!read_file(with_extension(a.input, "vsri"), text)) return fail("failed to read vsri");
    atsc_vsri *index = atsc_vsri_from_text((const char *)text.data(), text.size());
    if (!index) return fail("failed to read vsri");

    const std::string wbro_path = with_extension(output_base, "wbro");
    if (!write_wbro(wbro_path, data)) {
        atsc_vsri_free(index);
        return fail("cannot write " + wbro_path);
    }
    // Metric::get_samples (metric.rs:86-97): the i-th value gets vsri.get_time(i)
    std::vector<Sample> samples(data.size());
    for (size_t i = 0; i < data.size(); i++) {
        int32_t ts = 0;
        if (!atsc_vsri_get_time(index, (int32_t)i, &ts)) {
            atsc_vsri_free(index);
            return fail("vsri has no time for sample " + std::to_string(i));
        }
        samples[i] = {ts, data[i]};
    }
    atsc_vsri_free(index);
    if (!write_samples(with_extension(wbro_path, "csv"), samples)) return fail("failed to write samples to file");
    return 0;
}

// process_args, compress branch (main.rs:187-217)
int compress(atsc_ctx *ctx, const Args &a, const std::string &output_base) {
    std::vector<Sample> samples;
    std::string err;
    if (!read_samples(a.input, samples, err)) return fail("failed to read samples from file: " + err);
    // Metric::append_samples (metric.rs:53-65): one WavBrro + one VSRI, seconds within the day
    atsc_vsri *index = atsc_vsri_new();
    std::vector<double> values;
    values.reserve(samples.size());
    for (const Sample &s : samples) {
        if (atsc_vsri_update_for_point(index, atsc_day_elapsed_seconds(s.timestamp / 1000))) {
            atsc_vsri_free(index);
            return fail("failed to create metric from samples: updating for point failed, sample: Sample { timestamp: " +
                        std::to_string(s.timestamp) + ", value: " + ryu_f64(s.value) + " }");
        }
        values.push_back(s.value);
    }
    if (a.output_wavbrro && !write_wbro(with_extension(output_base, "wavbro"), values)) {
        atsc_vsri_free(index);
        return fail("cannot write wavbrro");
    }
    if (a.output_vsri) {
        std::string text(atsc_vsri_to_text(index, nullptr, 0), '\0');
        atsc_vsri_to_text(index, text.data(), text.size());
        if (!write_file(with_extension(output_base, "vsri"), (const uint8_t *)text.data(), text.size())) {
            atsc_vsri_free(index);
            return fail("failed to flush vsri to the file");
        }
    }
    atsc_vsri_free(index);
    if (a.no_compression) return 0;
    // compress_data (main.rs:97-132)
    std::vector<uint8_t> bro(values.size() * 16 + 4096 + 64 * (values.size() / 512 + 8));
    uint64_t off = 0, len = values.size(), boff = 0, blen = 0;
    double dummy = 0.0;
    int rc = atsc_gpu_compress_series(ctx, values.empty() ? &dummy : values.data(), &off, &len, 1, (uint8_t)a.compressor,
                                      a.error, a.speed, bro.data(), bro.size(), &boff, &blen, nullptr);
    if (rc) return fail(std::string("compress failed: ") + atsc_gpu_last_error(ctx));
    if (!write_file(with_extension(output_base, "bro"), bro.data() + boff, (size_t)blen))
        return fail("failed to write compressed data");
    return 0;
}

int usage() {
    fputs("A Time-Series compressor utilizes Brro Compressor for CSV format\n\n"
          "Usage: csv-compressor [OPTIONS] <INPUT>\n\nOptions:\n"
          "  -o, --output <OUTPUT>          Defines where the result will be stored\n"
          "  -u                             Defines if we should uncompress input\n"
          "      --no-compression           Disables compression operation\n"
          "      --output-vsri              Enables output of generated VSRI\n"
          "      --output-wavbrro           Enables output of generated WavBrro\n"
          "      --output-csv               Enable output result of decompression in CSV format\n"
          "      --compressor <COMPRESSOR>  [default: auto] [possible values: auto, noop, fft, constant, polynomial, idw]\n"
          "  -e, --error <ERROR>            maximum allowed error in percent, 0..50 [default: 5]\n"
          "  -c, --compression-selection-sample-level <N>  0..6 [default: 0]\n",
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
            static const char *names[] = {"noop", "fft", "idw", "constant", "polynomial", "auto"};
            int c = -1;
            for (int k = 0; k < 6; k++)
                if (v == names[k]) c = k;
            if (c < 0) return usage();
            a.compressor = c;
        } else if (s == "-e" || s == "--error" || s.rfind("--error=", 0) == 0) {
            a.error = (unsigned)atoi(value("--error").c_str());
            if (a.error > 50) return usage();
        } else if (s == "-o" || s == "--output" || s.rfind("--output=", 0) == 0)
            a.output = value("--output");
        else if (s == "-u")
            a.uncompress = true;
        else if (s == "--no-compression")
            a.no_compression = true;
        else if (s == "--output-vsri")
            a.output_vsri = true;
        else if (s == "--output-wavbrro")
            a.output_wavbrro = true;
        else if (s == "--output-csv")
            a.output_csv = true;  // accepted and unused, as in the reference (main.rs:62-64)
        else if (s == "-c" || s == "--compression-selection-sample-level" ||
                 s.rfind("--compression-selection-sample-level=", 0) == 0) {
            a.speed = (unsigned)atoi(value("--compression-selection-sample-level").c_str());
            if (a.speed > 6) return usage();
        } else if (s == "-h" || s == "--help")
            return usage();
        else if (!s.empty() && s[0] == '-')
            return usage();
        else
            a.input = s;
    }
    if (a.input.empty()) return usage();
    struct stat st;
    if (stat(a.input.c_str(), &st) != 0) return fail("Failed to retrieve metadata of " + a.input);
    if (!S_ISREG(st.st_mode)) return fail("Input is not a file");
    const std::string output_base = a.output.empty() ? a.input : a.output;
    // --no-compression needs no device: the index and WavBrro outputs are host work
    atsc_ctx *ctx = nullptr;
    if (a.uncompress || !a.no_compression) {
        int rc = atsc_gpu_create(nullptr, 0, &ctx);
        if (rc) {
            fprintf(stderr, "csv-compressor: no usable CUDA device (status %d); this build has no CPU path\n", rc);
            return 1;
        }
    }
    int ret = a.uncompress ? uncompress(ctx, a, output_base) : compress(ctx, a, output_base);
    if (ctx) atsc_gpu_destroy(ctx);
    return ret;
}

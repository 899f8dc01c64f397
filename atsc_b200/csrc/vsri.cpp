// vsri.cpp -- "Very Small Rolo Index" (SURVEY.md section 8f, row N3): the timestamp index the
// reference's second CLI (csv-compressor) stores beside a .bro file.  Host-only, no GPU needed.
//   reference: vsri/src/lib.rs  (struct :102-108, update_for_point :249-285, get_sample :312-328,
//   get_time :331-353, flush_to :442-462, load :466-497)
// A series of timestamps (seconds of the day) becomes a list of line segments
//   [sample rate m, first sample x0, first timestamp y0, number of samples]
// written as text:  min_ts \n max_ts \n m,x0,y0,n \n ...
// Integer semantics are Rust's i32: the divisions truncate toward zero like C++'s.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/atsc_gpu.h"

// i32 arithmetic of the reference's release build: wrapping add / sub / mul (a hostile index file can
// hold any numbers); the one quotient that would trap, INT32_MIN / -1, is defined as INT32_MIN
static inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static inline int32_t wmul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
static inline int32_t wdiv(int32_t a, int32_t b) { return (a == INT32_MIN && b == -1) ? INT32_MIN : a / b; }

struct atsc_vsri {
    int32_t min_ts = 0, max_ts = 0;
    std::vector<int32_t> seg;  // 4 per segment
    size_t n() const { return seg.size() / 4; }
    const int32_t *at(size_t i) const { return &seg[4 * i]; }
    int32_t *at(size_t i) { return &seg[4 * i]; }
    // lib.rs:299-304 current_segment
    void current(int32_t out[4]) const {
        if (seg.empty()) {
            out[0] = out[1] = out[2] = out[3] = 0;
            return;
        }
        memcpy(out, at(n() - 1), 16);
    }
    int32_t sample_count() const {  // lib.rs:368-371
        int32_t c[4];
        current(c);
        return wadd(c[3], c[1]);
    }
};

extern "C" {

atsc_vsri *atsc_vsri_new(void) { return new atsc_vsri(); }
void atsc_vsri_free(atsc_vsri *v) { delete v; }

// lib.rs:31-40 day_elapsed_seconds (chrono UTC; pre-1970 timestamps count back from midnight too)
int32_t atsc_day_elapsed_seconds(int64_t timestamp_sec) {
    int64_t r = timestamp_sec % 86400;
    if (r < 0) r += 86400;
    return (int32_t)r;
}

// lib.rs:249-285 update_for_point.  Returns 0, or 1 for a point in the past
// (Error::UpdateIndexForPointError; the index is left unchanged).
int atsc_vsri_update_for_point(atsc_vsri *v, int32_t y) {
    if (y < v->max_ts) return 1;
    v->max_ts = y;
    int32_t last[4];
    v->current(last);
    auto fake = [&]() {  // lib.rs:393-400 create_fake_segment
        int32_t s[4] = {0, wadd(last[1], last[3]), y, 1};
        v->seg.insert(v->seg.end(), s, s + 4);
    };
    if (v->seg.empty()) {
        v->min_ts = y;
        fake();
        return 0;
    }
    if (last[0] == 0) {
        // lib.rs:375-388 generate_segment: the second point of a segment fixes its rate
        int32_t *s = v->at(v->n() - 1);
        s[0] = wsub(y, last[2]);
        s[3] = 2;
        return 0;
    }
    // lib.rs:410-428 fits_segment: the point must be the next one on the line
    const int32_t b = wsub(last[2], wmul(last[0], last[1]));
    const int32_t x = wdiv(wsub(y, b), last[0]);
    if (x == wadd(last[3], last[1])) {
        int32_t *s = v->at(v->n() - 1);
        s[3] = wadd(s[3], 1);
        return 0;
    }
    fake();
    return 0;
}

int32_t atsc_vsri_min(const atsc_vsri *v) { return v->min_ts; }
int32_t atsc_vsri_max(const atsc_vsri *v) { return v->max_ts; }
int32_t atsc_vsri_sample_count(const atsc_vsri *v) { return v->sample_count(); }
uint64_t atsc_vsri_segment_count(const atsc_vsri *v) { return v->n(); }

// lib.rs:312-328 get_sample: 1 and *x when a sample exists exactly at time y
int atsc_vsri_get_sample(const atsc_vsri *v, int32_t y, int32_t *x) {
    for (size_t i = 0; i < v->n(); i++) {
        const int32_t *s = v->at(i);
        const int32_t end = wadd(s[2], wmul(s[0], wsub(s[3], 1)));
        if (y >= s[2] && y <= end) {
            if (s[0] == 0) return 0;  // a one-point segment: the reference divides by zero (panic)
            *x = wdiv(wsub(y, wsub(s[2], wmul(s[0], s[1]))), s[0]);
            return 1;
        }
    }
    return 0;
}

// lib.rs:331-353 get_time (the `segment[2] + segment[0] * x` of the reference is kept as is: it
// multiplies by the absolute sample number, not by the offset inside the segment)
int atsc_vsri_get_time(const atsc_vsri *v, int32_t x, int32_t *y) {
    if (x == 0) {
        *y = v->min_ts;
        return 1;
    }
    const int32_t cnt = v->sample_count();
    if (x > cnt) return 0;
    if (x == cnt) {
        *y = v->max_ts;
        return 1;
    }
    for (size_t i = 0; i < v->n(); i++) {
        const int32_t *s = v->at(i);
        if (x >= s[1] && x < wadd(s[1], s[3])) {
            *y = wadd(s[2], wmul(s[0], x));
            return 1;
        }
    }
    return 0;
}

// lib.rs:156-172 get_next_sample / :178-197 get_previous_sample: 1 and *x, or 0 for None
int atsc_vsri_get_next_sample(const atsc_vsri *v, int32_t y, int32_t *x) {
    if (y < v->min_ts) {
        *x = 0;
        return 1;
    }
    if (y >= v->max_ts) return 0;
    for (size_t i = v->n(); i-- > 0;) {
        const int32_t *s = v->at(i);
        if (y <= s[2]) {
            *x = s[1];
            return 1;
        }
    }
    return 0;
}
int atsc_vsri_get_previous_sample(const atsc_vsri *v, int32_t y, int32_t *x) {
    if (y < v->min_ts) return 0;
    if (y >= v->max_ts) {
        *x = v->sample_count();
        return 1;
    }
    for (size_t i = 0; i < v->n(); i++) {
        const int32_t *s = v->at(i);
        if (y < s[2]) {
            *x = wsub(s[1], 1);
            return 1;
        }
    }
    return 0;
}

// lib.rs:202-245 is_empty: does [t0, t1] fall between the indexed segments?
int atsc_vsri_is_empty(const atsc_vsri *v, int32_t t0, int32_t t1) {
    if (v->n() == 1) {
        if ((t0 >= v->min_ts && t0 <= v->max_ts) || (t1 <= v->max_ts && t1 >= v->min_ts)) return 0;
        if (t0 < v->min_ts && t1 > v->max_ts) return 0;
        return 1;
    }
    int32_t prev_end = 0;
    for (size_t i = 0; i < v->n(); i++) {
        const int32_t *s = v->at(i);
        const int32_t end = wadd(s[2], wmul(s[0], wsub(s[3], 1)));
        if (i >= 1 && t0 > prev_end && t1 < s[2]) return 1;
        if ((t0 >= s[2] && t0 < end) || (t1 < end && t1 >= s[2])) return 0;
        if (t0 < s[2] && t1 > end) return 0;
        prev_end = end;
    }
    return 1;
}

// lib.rs:356-366 get_all_timestamps: returns the count, writes min(count, cap) values
uint64_t atsc_vsri_all_timestamps(const atsc_vsri *v, int32_t *out, uint64_t cap) {
    uint64_t k = 0;
    for (size_t i = 0; i < v->n(); i++) {
        const int32_t *s = v->at(i);
        const uint64_t cnt = s[3] > 0 ? (uint64_t)s[3] : 0;
        const uint64_t room = (out && k < cap) ? cap - k : 0;  // only what fits is generated: a hostile count costs nothing
        for (uint64_t f = 0; f < cnt && f < room; f++) out[k + f] = wadd(wmul((int32_t)f, s[0]), s[2]);
        k += cnt;
    }
    return k;
}

// lib.rs:442-462 flush_to: the text image of the index; returns its length, writes it when it fits
uint64_t atsc_vsri_to_text(const atsc_vsri *v, char *out, uint64_t cap) {
    std::string t = std::to_string(v->min_ts) + "\n" + std::to_string(v->max_ts) + "\n";
    for (size_t i = 0; i < v->n(); i++) {
        const int32_t *s = v->at(i);
        t += std::to_string(s[0]) + "," + std::to_string(s[1]) + "," + std::to_string(s[2]) + "," + std::to_string(s[3]) + "\n";
    }
    if (out && t.size() <= cap) memcpy(out, t.data(), t.size());
    return t.size();
}

// lib.rs:466-497 load: nullptr where the reference's unwrap() would panic (bad integer, wrong field count)
atsc_vsri *atsc_vsri_from_text(const char *text, uint64_t len) {
    auto parse_i32 = [](std::string s, int32_t &o) {
        size_t a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r");
        if (a == std::string::npos) return false;
        s = s.substr(a, b - a + 1);
        if (!s.empty() && s[0] == '+') s = s.substr(1);  // Rust's i32::from_str accepts a leading '+'
        char *e = nullptr;
        long long x = strtoll(s.c_str(), &e, 10);
        if (e == s.c_str() || *e || x < INT32_MIN || x > INT32_MAX) return false;
        o = (int32_t)x;
        return true;
    };
    atsc_vsri *v = new atsc_vsri();
    uint64_t pos = 0;
    int line_no = 1;
    while (pos < len) {
        uint64_t e = pos;
        while (e < len && text[e] != '\n') e++;
        std::string line(text + pos, text + e);
        pos = e + 1;
        bool ok = true;
        if (line_no == 1)
            ok = parse_i32(line, v->min_ts);
        else if (line_no == 2)
            ok = parse_i32(line, v->max_ts);
        else {
            int32_t s[4];
            int k = 0;
            size_t p = 0;
            while (ok) {
                size_t c = line.find(',', p);
                std::string f = line.substr(p, c == std::string::npos ? std::string::npos : c - p);
                if (k >= 4 || !parse_i32(f, s[k])) ok = false;
                k++;
                if (c == std::string::npos) break;
                p = c + 1;
            }
            if (ok && k == 4)
                v->seg.insert(v->seg.end(), s, s + 4);
            else
                ok = false;
        }
        if (!ok) {
            delete v;
            return nullptr;
        }
        line_no++;
    }
    return v;
}

}  // extern "C"

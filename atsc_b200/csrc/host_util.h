// host_util.h -- host-side integer helpers shared by api.cu and the stream layer.
// Reference citations are relative to /root/reference/atsc/src/.
#pragma once
#include <stdint.h>

#include <vector>

namespace atsc_host {

// utils/mod.rs:41-49
inline bool is_decomposable(uint64_t n) {
    if (n == 0) return false;
    while (n % 2 == 0) n /= 2;
    while (n % 3 == 0) n /= 3;
    return n == 1;
}
// utils/mod.rs:32-38: smallest 2^a*3^b strictly greater than n
inline uint64_t next_size(uint64_t n) {
    n += 1;
    while (!is_decomposable(n)) n += 1;
    return n;
}
// utils/mod.rs:24-29
inline uint64_t prev_power_of_two(uint64_t n) {
    int hb = 63 - __builtin_clzll(n | 1);
    return (1ull << hb) & n;
}
// optimizer/mod.rs:78-98 get_chunks_sizes (MAX_FRAME_SIZE 131072, MIN_FRAME_SIZE 512)
inline void chunk_sizes(uint64_t len, std::vector<uint32_t> &out) {
    while (len > 0) {
        uint64_t s;
        if (len >= 131072)
            s = 131072;
        else if (len <= 512)
            s = len;
        else
            s = prev_power_of_two(len);
        out.push_back((uint32_t)s);
        len -= s;
    }
}

// bincode 2 "standard" varint (compressor/mod.rs:126-130)
inline void put_varint(std::vector<uint8_t> &b, uint64_t u) {
    if (u < 251) {
        b.push_back((uint8_t)u);
    } else if (u < 65536ull) {
        b.push_back(251);
        for (int i = 0; i < 2; i++) b.push_back((uint8_t)(u >> (8 * i)));
    } else if (u < 4294967296ull) {
        b.push_back(252);
        for (int i = 0; i < 4; i++) b.push_back((uint8_t)(u >> (8 * i)));
    } else {
        b.push_back(253);
        for (int i = 0; i < 8; i++) b.push_back((uint8_t)(u >> (8 * i)));
    }
}
// returns false on truncated / invalid input
inline bool get_varint(const uint8_t *p, uint64_t len, uint64_t &pos, uint64_t &v) {
    if (pos >= len) return false;
    uint8_t t = p[pos++];
    if (t < 251) {
        v = t;
        return true;
    }
    int nb = t == 251 ? 2 : t == 252 ? 4 : t == 253 ? 8 : -1;
    if (nb < 0 || pos + nb > len) return false;
    v = 0;
    for (int i = 0; i < nb; i++) v |= (uint64_t)p[pos + i] << (8 * i);
    pos += nb;
    return true;
}

}  // namespace atsc_host

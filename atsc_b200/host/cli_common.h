// cli_common.h -- file helpers shared by the two command lines (atsc, csv-compressor).
#pragma once
#include <cstdint>
#include <fstream>
#include <string>
#include <vector>

namespace {

bool read_file(const std::string &path, std::vector<uint8_t> &out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    f.seekg(0, std::ios::end);
    std::streamsize n = f.tellg();
    f.seekg(0);
    out.resize((size_t)n);
    return n == 0 || (bool)f.read((char *)out.data(), n);
}
bool write_file(const std::string &path, const uint8_t *p, size_t n) {
    std::ofstream f(path, std::ios::binary | std::ios::trunc);
    if (!f) return false;
    f.write((const char *)p, (std::streamsize)n);
    return (bool)f;
}
// PathBuf::set_extension
std::string with_extension(const std::string &path, const char *ext) {
    size_t slash = path.find_last_of('/');
    size_t dot = path.find_last_of('.');
    std::string stem = (dot != std::string::npos && (slash == std::string::npos || dot > slash + 1)) ? path.substr(0, dot) : path;
    return stem + "." + ext;
}

}  // namespace

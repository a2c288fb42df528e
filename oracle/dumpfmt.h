// TEST INFRASTRUCTURE (oracle/): tiny tagged-array dump format shared by the two
// reference harnesses and read back by tests/_dumpfmt.py.
//   record := u32 name_len | name bytes | u8 dtype ('f','d','i','B','H') | u32 ndim | u64 dims[ndim] | raw data
#pragma once
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

struct DumpWriter {
    FILE* f = nullptr;
    explicit DumpWriter(const char* path) { f = fopen(path, "wb"); }
    ~DumpWriter() { if (f) fclose(f); }
    bool ok() const { return f != nullptr; }
    void put(const std::string& name, char dtype, const std::vector<uint64_t>& dims, const void* data, size_t elem) {
        uint32_t nl = (uint32_t)name.size();
        fwrite(&nl, 4, 1, f); fwrite(name.data(), 1, nl, f);
        fwrite(&dtype, 1, 1, f);
        uint32_t nd = (uint32_t)dims.size(); fwrite(&nd, 4, 1, f);
        uint64_t n = 1;
        for (uint64_t d : dims) { fwrite(&d, 8, 1, f); n *= d; }
        if (n) fwrite(data, elem, n, f);
    }
    void f32(const std::string& n, const float* p, std::vector<uint64_t> d) { put(n, 'f', d, p, 4); }
    void f64(const std::string& n, const double* p, std::vector<uint64_t> d) { put(n, 'd', d, p, 8); }
    void i32(const std::string& n, const int32_t* p, std::vector<uint64_t> d) { put(n, 'i', d, p, 4); }
    void u8(const std::string& n, const uint8_t* p, std::vector<uint64_t> d) { put(n, 'B', d, p, 1); }
    void u16(const std::string& n, const uint16_t* p, std::vector<uint64_t> d) { put(n, 'H', d, p, 2); }
    void scalar_i(const std::string& n, int32_t v) { i32(n, &v, {1}); }
    void scalar_f(const std::string& n, float v) { f32(n, &v, {1}); }
};

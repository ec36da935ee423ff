// .b2w weight container (written by whisper.coreml_b200/export.py) -> one device allocation.
//
//   char magic[4] = "B2W1"; uint32 n_tensors; uint64 data_offset; uint64 data_bytes;
//   n_tensors x { char name[64]; uint32 dtype (0 f32, 1 bf16, 2 i32); uint32 ndim; uint64 shape[4];
//                 uint64 offset (from data_offset, 256-byte aligned); uint64 nbytes; }
#pragma once
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace b200 {

struct TensorInfo {
    int dtype = 0, ndim = 0;
    long shape[4] = {0, 0, 0, 0};
    size_t nbytes = 0;
    void* dev = nullptr;
    const void* host = nullptr;     // valid until drop_host()
};

class WeightFile {
public:
    bool load(const std::string& path);      // false + record_error on failure
    void unload();
    bool loaded() const { return base_ != nullptr; }
    const TensorInfo* find(const std::string& name) const;
    // typed getters: record an error and return nullptr when missing
    const float* f32(const std::string& name) const;
    const bf16* b16(const std::string& name) const;
    const int* i32_host(const std::string& name) const;   // host copy (small metadata tensors)
    size_t device_bytes() const { return bytes_; }
    const std::string& path() const { return path_; }
    int refcount = 0;

private:
    std::string path_;
    std::map<std::string, TensorInfo> tensors_;
    std::vector<char> host_;
    void* base_ = nullptr;
    size_t bytes_ = 0;
};

}  // namespace b200

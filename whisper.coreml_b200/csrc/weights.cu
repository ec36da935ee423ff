#include "weights.cuh"

#include <string.h>

namespace b200 {

namespace {
struct FileHeader { char magic[4]; uint32_t n_tensors; uint64_t data_offset; uint64_t data_bytes; };
struct FileEntry { char name[64]; uint32_t dtype; uint32_t ndim; uint64_t shape[4]; uint64_t offset; uint64_t nbytes; };
}  // namespace

bool WeightFile::load(const std::string& path) {
    if (loaded()) return true;
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { record_error("cannot open weight file %s", path.c_str()); return false; }
    FileHeader h;
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "B2W1", 4) != 0) {
        record_error("%s is not a B2W1 container", path.c_str()); fclose(f); return false;
    }
    std::vector<FileEntry> entries(h.n_tensors);
    if (fread(entries.data(), sizeof(FileEntry), h.n_tensors, f) != h.n_tensors) {
        record_error("%s: truncated index", path.c_str()); fclose(f); return false;
    }
    host_.resize(h.data_bytes);
    fseek(f, (long)h.data_offset, SEEK_SET);
    if (fread(host_.data(), 1, h.data_bytes, f) != h.data_bytes) {
        record_error("%s: truncated data", path.c_str()); fclose(f); host_.clear(); return false;
    }
    fclose(f);
    cudaError_t e = cudaMalloc(&base_, h.data_bytes ? h.data_bytes : 256);
    if (e != cudaSuccess) { record_error("cudaMalloc(%zu) for %s: %s", (size_t)h.data_bytes, path.c_str(), cudaGetErrorString(e)); base_ = nullptr; return false; }
    B200_CHECK(cudaMemcpy(base_, host_.data(), h.data_bytes, cudaMemcpyHostToDevice));
    bytes_ = h.data_bytes;
    for (const FileEntry& en : entries) {
        TensorInfo t;
        t.dtype = (int)en.dtype; t.ndim = (int)en.ndim;
        for (int i = 0; i < 4; ++i) t.shape[i] = (long)en.shape[i];
        t.nbytes = en.nbytes;
        t.dev = (char*)base_ + en.offset;
        t.host = host_.data() + en.offset;
        char nm[65]; memcpy(nm, en.name, 64); nm[64] = 0;
        tensors_[nm] = t;
    }
    path_ = path;
    return true;
}

void WeightFile::unload() {
    if (base_) B200_CHECK(cudaFree(base_));
    base_ = nullptr; bytes_ = 0; tensors_.clear(); host_.clear(); host_.shrink_to_fit(); path_.clear();
}

const TensorInfo* WeightFile::find(const std::string& name) const {
    auto it = tensors_.find(name);
    return it == tensors_.end() ? nullptr : &it->second;
}
const float* WeightFile::f32(const std::string& name) const {
    const TensorInfo* t = find(name);
    if (!t || t->dtype != 0) { record_error("weight '%s' (f32) missing in %s", name.c_str(), path_.c_str()); return nullptr; }
    return (const float*)t->dev;
}
const bf16* WeightFile::b16(const std::string& name) const {
    const TensorInfo* t = find(name);
    if (!t || t->dtype != 1) { record_error("weight '%s' (bf16) missing in %s", name.c_str(), path_.c_str()); return nullptr; }
    return (const bf16*)t->dev;
}
const int* WeightFile::i32_host(const std::string& name) const {
    const TensorInfo* t = find(name);
    if (!t || t->dtype != 2) { record_error("tensor '%s' (i32) missing in %s", name.c_str(), path_.c_str()); return nullptr; }
    return (const int*)t->host;
}

}  // namespace b200

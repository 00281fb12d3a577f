// weights_file.cpp -- see weights_file.h
#include "weights_file.h"

#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace pkb {

namespace {
#pragma pack(push, 1)
struct Header { char magic[8]; uint32_t version, n_tensors, n_cfg, pad; };
struct CfgEntry { char key[32]; int64_t value; };
struct TensorEntry { char name[96]; uint32_t dtype, ndim; uint64_t dims[4]; uint64_t offset, nbytes; };
#pragma pack(pop)
}  // namespace

WeightsFile::WeightsFile(const std::string& path) {
  int fd = ::open(path.c_str(), O_RDONLY);
  if (fd < 0) throw std::runtime_error("cannot open weights file: " + path);
  struct stat sb;
  if (fstat(fd, &sb) != 0 || (size_t)sb.st_size < sizeof(Header)) {
    ::close(fd);
    throw std::runtime_error("weights file too small: " + path);
  }
  size_ = (size_t)sb.st_size;
  map_ = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
  ::close(fd);
  if (map_ == MAP_FAILED) { map_ = nullptr; throw std::runtime_error("mmap failed: " + path); }
  try {
    parse(path);
  } catch (...) {      // a constructor that throws does not run the destructor: release the mapping here
    munmap(map_, size_);
    map_ = nullptr;
    throw;
  }
}

void WeightsFile::parse(const std::string& path) {
  const unsigned char* p = static_cast<const unsigned char*>(map_);
  Header h;
  memcpy(&h, p, sizeof(h));
  if (memcmp(h.magic, "PKB200W1", 8) != 0 || h.version != 1) throw std::runtime_error("not a PKB200W1 weights file: " + path);
  size_t pos = sizeof(Header);
  if (pos + (size_t)h.n_cfg * sizeof(CfgEntry) + (size_t)h.n_tensors * sizeof(TensorEntry) > size_)
    throw std::runtime_error("weights file truncated (tables): " + path);
  for (uint32_t i = 0; i < h.n_cfg; ++i, pos += sizeof(CfgEntry)) {
    CfgEntry c;
    memcpy(&c, p + pos, sizeof(c));
    cfg_[std::string(c.key, strnlen(c.key, sizeof(c.key)))] = c.value;
  }
  for (uint32_t i = 0; i < h.n_tensors; ++i, pos += sizeof(TensorEntry)) {
    TensorEntry t;
    memcpy(&t, p + pos, sizeof(t));
    if (t.ndim > 4 || t.offset + t.nbytes > size_) throw std::runtime_error("weights file truncated (data): " + path);
    HostTensor ht;
    ht.dtype = (int)t.dtype;
    ht.dims.assign(t.dims, t.dims + t.ndim);
    ht.data = p + t.offset;
    ht.nbytes = t.nbytes;
    if (ht.numel() * (ht.dtype == 1 ? 2 : 4) != ht.nbytes) throw std::runtime_error("weights file: size mismatch in a tensor entry");
    tensors_[std::string(t.name, strnlen(t.name, sizeof(t.name)))] = ht;
  }
}

WeightsFile::~WeightsFile() {
  if (map_) munmap(map_, size_);
}

int64_t WeightsFile::cfg(const std::string& key) const {
  auto it = cfg_.find(key);
  if (it == cfg_.end()) throw std::runtime_error("weights file: missing config key " + key);
  return it->second;
}
int64_t WeightsFile::cfg_or(const std::string& key, int64_t dflt) const {
  auto it = cfg_.find(key);
  return it == cfg_.end() ? dflt : it->second;
}
const HostTensor& WeightsFile::get(const std::string& name) const {
  auto it = tensors_.find(name);
  if (it == tensors_.end()) throw std::runtime_error("weights file: missing tensor " + name);
  return it->second;
}
std::vector<float> WeightsFile::f32(const std::string& name) const {
  const HostTensor& t = get(name);
  std::vector<float> out(t.numel());
  if (t.dtype == 0) {
    memcpy(out.data(), t.data, t.nbytes);
  } else {
    const uint16_t* s = reinterpret_cast<const uint16_t*>(t.data);
    for (size_t i = 0; i < out.size(); ++i) out[i] = bf16_bits_to_f32(s[i]);
  }
  return out;
}
std::vector<uint16_t> WeightsFile::bf16(const std::string& name) const {
  const HostTensor& t = get(name);
  std::vector<uint16_t> out(t.numel());
  if (t.dtype == 1) {
    memcpy(out.data(), t.data, t.nbytes);
  } else {
    const float* s = reinterpret_cast<const float*>(t.data);
    for (size_t i = 0; i < out.size(); ++i) out[i] = f32_to_bf16_bits(s[i]);
  }
  return out;
}

}  // namespace pkb

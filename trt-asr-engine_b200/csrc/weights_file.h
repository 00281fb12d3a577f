// weights_file.h -- reader for <model_dir>/weights.bin (format: ../weights_io.py).  Host only.
//
// Replaces the engine deserialisation of the reference (/root/reference/cpp/src/parakeet_trt.cpp:1712-1738: three
// TensorRT plan files) by one flat container of NeMo state_dict tensors.
#pragma once
#include <stdint.h>

#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace pkb {

struct HostTensor {
  int dtype = 0;  // 0 = f32, 1 = bf16
  std::vector<uint64_t> dims;
  const unsigned char* data = nullptr;
  uint64_t nbytes = 0;
  uint64_t numel() const {
    uint64_t n = 1;
    for (auto d : dims) n *= d;
    return n;
  }
};

class WeightsFile {
 public:
  explicit WeightsFile(const std::string& path);
  ~WeightsFile();
  WeightsFile(const WeightsFile&) = delete;
  int64_t cfg(const std::string& key) const;
  int64_t cfg_or(const std::string& key, int64_t dflt) const;
  const HostTensor& get(const std::string& name) const;
  bool has(const std::string& name) const { return tensors_.count(name) != 0; }
  // convenience: tensor as f32 vector (bf16 widened exactly)
  std::vector<float> f32(const std::string& name) const;
  // raw bf16 bit patterns (tensor must be stored as bf16)
  std::vector<uint16_t> bf16(const std::string& name) const;

 private:
  void parse(const std::string& path);
  void* map_ = nullptr;
  size_t size_ = 0;
  std::map<std::string, int64_t> cfg_;
  std::map<std::string, HostTensor> tensors_;
};

inline uint16_t f32_to_bf16_bits(float f) {
  uint32_t u;
  __builtin_memcpy(&u, &f, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
inline float bf16_bits_to_f32(uint16_t b) {
  uint32_t u = ((uint32_t)b) << 16;
  float f;
  __builtin_memcpy(&f, &u, 4);
  return f;
}

}  // namespace pkb

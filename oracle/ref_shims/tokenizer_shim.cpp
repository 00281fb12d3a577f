// oracle/ref_shims/tokenizer_shim.cpp -- TEST INFRASTRUCTURE.
// extern "C" wrapper so tests can call the REFERENCE's own Tokenizer (compiled from
// /root/reference/cpp/src/tokenizer.cpp where it lies; see oracle/Makefile) through ctypes.
#include "tokenizer.h"

#include <cstring>
#include <string>
#include <vector>

extern "C" {
void* reftok_open(const char* vocab_path) {
  try { return new Tokenizer(vocab_path); } catch (...) { return nullptr; }
}
void reftok_close(void* t) { delete static_cast<Tokenizer*>(t); }
int reftok_decode(void* t, const int* ids, int n, char* out, int cap) {
  std::string s = static_cast<Tokenizer*>(t)->decode(std::vector<int>(ids, ids + n));
  if ((int)s.size() + 1 > cap) return -1;
  std::memcpy(out, s.c_str(), s.size() + 1);
  return (int)s.size();
}
int reftok_is_punct_only(void* t, int id) { return static_cast<Tokenizer*>(t)->is_punct_only(id) ? 1 : 0; }
int reftok_vocab_size(void* t) { return static_cast<Tokenizer*>(t)->vocab_size(); }
}

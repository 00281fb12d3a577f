// vocab.cpp -- see vocab.h.
#include "vocab.h"

#include <cctype>
#include <fstream>
#include <stdexcept>

namespace pkb {

namespace {
constexpr unsigned char kMarker[3] = {0xE2, 0x96, 0x81};      // U+2581, SentencePiece's word-start marker
bool has_marker(const std::string& s) {
  return s.size() >= 3 && (unsigned char)s[0] == kMarker[0] && (unsigned char)s[1] == kMarker[1] && (unsigned char)s[2] == kMarker[2];
}
}  // namespace

Vocab::Vocab(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("cannot open " + path);
  std::string line;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    pieces_.push_back(line);
  }
  if (pieces_.empty()) throw std::runtime_error("vocab file is empty: " + path);
  flags_.resize(pieces_.size(), 0);
  for (size_t i = 0; i < pieces_.size(); ++i) {
    const std::string& p = pieces_[i];
    uint8_t fl = 0;
    if (p == "<blank>" || p == "<pad>" || p == "<unk>" || (!p.empty() && p.front() == '<' && p.back() == '>')) fl |= kSpecial;
    const bool marker = has_marker(p);
    if (marker) fl |= kWordStart;
    if (!(fl & kSpecial)) {
      // punctuation-only: after the optional marker at least one byte, none alphanumeric, at least one not white space
      bool alnum = false, non_space = false;
      for (size_t k = marker ? 3 : 0; k < p.size(); ++k) {
        const unsigned char c = (unsigned char)p[k];
        if (std::isalnum(c)) { alnum = true; break; }
        if (!std::isspace(c)) non_space = true;
      }
      if (non_space && !alnum) fl |= kPunctOnly;
    }
    flags_[i] = fl;
  }
}

const std::string& Vocab::piece(int id) const {
  static const std::string empty;
  return id >= 0 && id < size() ? pieces_[id] : empty;
}

int Vocab::find(const std::string& piece) const {
  for (size_t i = 0; i < pieces_.size(); ++i)
    if (pieces_[i] == piece) return (int)i;
  return -1;
}

std::string Vocab::decode(const int* ids, size_t n) const {
  std::string out;
  for (size_t k = 0; k < n; ++k) {
    const int id = ids[k];
    if (id < 0 || id >= size() || (flags_[id] & kSpecial)) continue;
    const std::string& p = pieces_[id];
    if (flags_[id] & kWordStart) {
      if (!out.empty() && out.back() != ' ') out.push_back(' ');
      out.append(p, 3, std::string::npos);
    } else {
      out.append(p);
    }
  }
  const size_t first = out.find_first_not_of(' ');
  return first == std::string::npos ? std::string() : out.substr(first);
}

std::vector<uint32_t> Vocab::punct_bitmap(int n_bits) const {
  std::vector<uint32_t> bits((size_t)(n_bits + 31) / 32, 0u);
  for (int i = 0; i < size() && i < n_bits; ++i)
    if (flags_[i] & kPunctOnly) bits[i >> 5] |= 1u << (i & 31);
  return bits;
}

}  // namespace pkb
